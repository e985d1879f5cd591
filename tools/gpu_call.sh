set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "custom_grids or throughput or golden" > gpurun_out/g12_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g12_tests.log
tail -6 gpurun_out/g12_tests.log
B="python bench.py --no-cpu --no-e2e --no-pipeline --no-configs --steps 5"
run() { python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$1',d['value'],d['roofline']['kernel_ms_per_launch'], d['fit'].get('slow_path_voxels'))"; }
$B 2>/dev/null | run base
AMX_STAGE3_WARPS=32 AMX_CAP_STAGE3=24 $B 2>/dev/null | run s3w32cap24
AMX_STAGE3_WARPS=32 AMX_CAP_STAGE3=20 $B 2>/dev/null | run s3w32cap20
AMX_CAP_STAGE3=24 $B 2>/dev/null | run s3cap24
AMX_CAP_STAGE1=12 $B 2>/dev/null | run s1cap12
AMX_CAP_STAGE2=28 $B 2>/dev/null | run s2cap28
python - <<'PY'
import torch, numpy as np, sys, time
sys.path.insert(0,'.')
from amico_b200 import synth
from amico_b200.plan import Plan
for cfg,model,n in ((2,'NODDI',20000),(1,'FreeWater',200000),(5,'CylinderZeppelinBall',100000),(4,'SANDI',500000)):
    P=synth.make_problem(cfg,n_vox=n,model=model)
    plan=Plan(model,P.KERNELS,P.htable,P.params,dwi_idx=P.scheme.dwi_idx)
    y=torch.from_numpy(P.y).cuda(); d=None if model=='SANDI' else torch.from_numpy(np.array(P.DIRs)).cuda()
    from oracle import oracle as orc
    l1,l2=orc.DEFAULT_LAMBDAS[model]
    for ex in (True,):
        plan.fit(y,d,l1,l2,exact=ex); torch.cuda.synchronize(); t=time.time(); plan.fit(y,d,l1,l2,exact=ex); torch.cuda.synchronize(); dt=time.time()-t
        print(model,'exact' if ex else 'fast', n/dt, 'voxels/s', plan.last_counters())
PY
