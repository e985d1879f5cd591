set -x
python tools/nan_probe3.py 23 2>&1 | grep '"cfg": 3' | sed 's/"timing".*"counters"/"counters"/' | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ordered_after or second_generation or at_scale" 2>&1 | tail -3
bash tools/bench_sweep.sh "AMX_LEAN2=1" 2>&1 | tail -1
