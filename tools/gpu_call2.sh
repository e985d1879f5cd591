set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_r02_2gpu_b.json 2> gpurun_out/bench_r02_2gpu_b.err; tail -2 gpurun_out/bench_r02_2gpu_b.err
python -c "
import json
for l in open('gpurun_out/bench_r02_2gpu_b.json'):
    if l.startswith('{'): d=json.loads(l); print(d['value']); print(d['sharded_cfg3'])
"
