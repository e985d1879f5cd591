"""Probe: fit the rank-1 slab of the sharded cfg3 bench on one GPU and report non-finite map values per kernel variant."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amico_b200 import synth, models as amx_models
from amico_b200.plan import Plan
dev = torch.device("cuda:0")
cfg = 3
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5242880
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20251017 + cfg + 1000
P = synth.make_problem(cfg, n_vox=8)
mdl = amx_models.NODDI(); mdl.set_solver()
l1, l2 = mdl.solver_params["lambda1"], mdl.solver_params["lambda2"]
if os.environ.get("PROBE_POISON"):  # hand the driver back memory full of NaN bit patterns: the plan's cudaMalloc's will recycle it
    junk = [torch.full((1 << 28,), float("nan"), dtype=torch.float64, device=dev) for _ in range(int(os.environ["PROBE_POISON"]))]
    torch.cuda.synchronize(); del junk; torch.cuda.empty_cache()
plan = Plan("NODDI", P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx, device=0)
y, d = synth.make_voxels_torch("NODDI", P.KERNELS, P.htable, n, seed, dev)
est = torch.full((n, 3), float("nan"), dtype=torch.float64, device=dev)
plan.fit(y, d.clone(), l1, l2, out=est)
torch.cuda.synchronize()
bad = ~torch.isfinite(est).all(dim=1)
idx = torch.nonzero(bad).flatten()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("AMX_")}, "n": n, "bad_voxels": int(bad.sum().item()), "counters": plan.last_counters(),
                  "first": idx[:8].tolist(), "y_min": float(y.min().item()), "y_nonfinite": int((~torch.isfinite(y)).sum().item())}))
if len(idx):
    i = int(idx[0]); print("voxel", i, "est", est[i].tolist(), "y[:8]", y[i, :8].tolist(), "dir", d[i].tolist())
