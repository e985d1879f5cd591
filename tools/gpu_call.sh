set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/g5_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g5_tests.log
tail -15 gpurun_out/g5_tests.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "throughput_kernel" 2>&1 | grep -E "pass fraction|passed|failed"
python tools/bench_models.py 1048576 FreeWater1,SANDI4,CylinderZeppelinBall5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('fast', d['model'], d['voxels_per_s'], d['ms'], d['counters']['warps_per_cta'], d['counters']['smem_bytes'])
    except Exception: print(l[:300])"
AMX_CZB_DENSE=0 python tools/bench_models.py 1048576 CylinderZeppelinBall5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('lars only', d['model'], d['voxels_per_s'], d['ms'])
    except Exception: print(l[:300])"
python bench.py --steps 3 --no-pipeline > gpurun_out/g5_bench.json 2> gpurun_out/g5_bench.err; tail -3 gpurun_out/g5_bench.err; python -c "
import json;d=json.load(open('gpurun_out/g5_bench.json'))
print(json.dumps({k:d[k] for k in ('value','e2e','e2e_plugin')},indent=0)[:1500])
for c,r in d['configs'].items(): print(c, r.get('voxels_per_s'), r.get('roofline_frac'), r.get('parity_vs_cpu'), r.get('error'))"
