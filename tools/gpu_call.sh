set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "noddi or slow_path" > gpurun_out/h14_tests.log 2>&1; tail -25 gpurun_out/h14_tests.log
bash tools/bench_sweep.sh "AMX_LEAN2=1" > gpurun_out/h14_sweep.log 2>&1
cat gpurun_out/h14_sweep.log
