for caps in "32 32" "24 24" "20 24" "16 20" "24 32"; do set -- $caps; AMX_CAP_STAGE2=$1 AMX_CAP_STAGE3=$2 python - <<PY
import json,subprocess,sys,os
import numpy as np, torch
sys.path.insert(0,'.')
from amico_b200 import synth
from amico_b200.plan import Plan
P = synth.make_problem(2)
plan = Plan('NODDI', P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx)
y = torch.from_numpy(P.y).cuda(); d = torch.from_numpy(np.ascontiguousarray(P.DIRs)).cuda()
for _ in range(3): r = plan.fit(y, d.clone(), 0.5, 1e-3)
t = plan.last_timing(); c = plan.last_counters()
print('caps', os.environ['AMX_CAP_STAGE2'], os.environ['AMX_CAP_STAGE3'], 'fit_kernel_ms %.2f' % t['fit_kernel_ms'], 'slow', c['slow_path_voxels'], 'smem', c['smem_bytes'], 'sum %.6f' % float(r['estimates'].sum()))
PY
done
