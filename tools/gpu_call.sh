set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/g11_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g11_tests.log
tail -8 gpurun_out/g11_tests.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "throughput_kernel" 2>&1 | grep -E "pass fraction|passed|failed"
python tools/bench_models.py 1048576 FreeWater1,SANDI4,CylinderZeppelinBall5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('fast', d['model'], d['voxels_per_s'], d['ms'])
    except Exception: print(l[:300])"
python bench.py --steps 5 --no-pipeline --no-cpu > gpurun_out/g11_bench.json 2> gpurun_out/g11_bench.err; tail -3 gpurun_out/g11_bench.err; python -c "
import json;d=json.load(open('gpurun_out/g11_bench.json'))
print('value',d['value'],d['roofline']['kernel_ms_per_launch'],'e2e',d['e2e']['value'],'plugin',d['e2e_plugin']['value'])
for c,r in d['configs'].items(): print(c, r.get('voxels_per_s'), r.get('roofline_frac'), 'nan',r.get('nan_values'),'exact',r.get('exact_path_voxels'),'slow',r.get('slow_path_voxels'), r.get('error'))"
python - <<'PY'
import torch, numpy as np, sys
sys.path.insert(0,'.')
from amico_b200 import synth, models
from amico_b200.plan import Plan
P=synth.make_problem(3,n_vox=8)
dev=torch.device('cuda',0)
y,d=synth.make_voxels_torch('NODDI',P.KERNELS,P.htable,10485760//2,20251017+3,dev)
plan=Plan('NODDI',P.KERNELS,P.htable,P.params,dwi_idx=P.scheme.dwi_idx)
est=plan.fit(y,d,0.5,1e-3)['estimates']
bad=torch.isnan(est).any(1).nonzero().flatten()
print('nan voxels',bad.numel(), bad[:10].tolist(), plan.last_counters())
if bad.numel():
    i=int(bad[0]); print('y',y[i][:10].tolist(),'ymax',float(y[i].max()),'d',d[i].tolist(), 'y nan', bool(torch.isnan(y[i]).any()))
    print('ynan total', int(torch.isnan(y).any(1).sum()), 'dnan', int(torch.isnan(d).any(1).sum()))
PY
