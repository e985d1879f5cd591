set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
bash tools/bench_sweep.sh "AMX_LEAN2=1" 2>&1 | tail -1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-pipeline --no-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('checksum', repr(d['maps_checksum']))"
python tools/bench_models.py 1048576 2>&1 | tail -3; python -c "
import json; d=json.load(open('gpurun_out/bench_models.json'))
for k,v in d.items(): print(k, round(v['voxels_per_s']/1e6,1))"
