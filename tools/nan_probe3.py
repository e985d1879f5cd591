"""Probe: a cfg2 plan is used and closed, then a cfg3 plan fits a 5.2M-voxel slab in the same process (what bench.py does)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amico_b200 import synth, models as amx_models
from amico_b200.plan import Plan
dev = torch.device("cuda:0")
mdl = amx_models.NODDI(); mdl.set_solver()
l1, l2 = mdl.solver_params["lambda1"], mdl.solver_params["lambda2"]
def run(cfg, n, seed, tag):
    P = synth.make_problem(cfg, n_vox=8)
    plan = Plan("NODDI", P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx, device=0)
    y, d = synth.make_voxels_torch("NODDI", P.KERNELS, P.htable, n, seed, dev)
    est = torch.full((n, 3), 777.0, dtype=torch.float64, device=dev)
    for it in range(3):
        plan.fit(y, d, l1, l2, out=est)
        torch.cuda.synchronize()
        bad = ~(torch.isfinite(est).all(dim=1) & (est != 777.0).all(dim=1))
        print(json.dumps({"tag": tag, "cfg": cfg, "n": n, "iter": it, "bad_voxels": int(bad.sum().item()), "timing": plan.last_timing(), "nan_rows_head": int((~torch.isfinite(est[:1000]).all(dim=1)).sum().item()), "counters": plan.last_counters()}), flush=True)
        if bad.any():
            idx = torch.nonzero(bad).flatten(); i = int(idx[0])
            print("  first bad", idx[:6].tolist(), "est", est[i].tolist(), flush=True)
    plan.close()
    del y, d, est
    torch.cuda.empty_cache()
order = sys.argv[1] if len(sys.argv) > 1 else "23"
N2 = int(sys.argv[2]) if len(sys.argv) > 2 else 1048576
N3 = int(sys.argv[3]) if len(sys.argv) > 3 else 5242880
for c in order:
    if c == "2": run(2, N2, 20251019, "cfg2")
    if c == "3": run(3, N3, 20251020, "cfg3")
