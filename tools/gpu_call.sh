set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "second_generation or known or at_scale or variants" 2>&1 | tail -2
bash tools/bench_sweep.sh "AMX_LEAN2=1" 2>&1 | cut -c1-60
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-pipeline --no-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('checksum', repr(d['maps_checksum']), d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'])"
