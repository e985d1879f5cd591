set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err; tail -2 gpurun_out/bench_r02_2gpu.err
