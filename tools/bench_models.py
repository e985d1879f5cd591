"""Device-resident throughput of every model on its benchmark protocol (BASELINE.json configs), N voxels each.
usage (GPU box): python tools/bench_models.py [n_vox]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amico_b200 import models as amx_models, synth  # noqa: E402
from amico_b200.plan import Plan  # noqa: E402

n_vox = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
out = {}
for cfg, model in ((1, "FreeWater"), (4, "SANDI"), (5, "CylinderZeppelinBall"), (3, "NODDI"), (2, "NODDI")):
    if only and f"{model}{cfg}" not in only:
        continue
    t0 = time.time()
    P = synth.make_problem(cfg, n_vox=n_vox, model=model)
    mdl = getattr(amx_models, model)()
    mdl.set_solver()
    l1, l2 = mdl.solver_params["lambda1"], mdl.solver_params["lambda2"]
    plan = Plan(model, P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx)
    y = torch.from_numpy(P.y).cuda()
    d = None if model == "SANDI" else torch.from_numpy(np.array(P.DIRs)).cuda()
    est = torch.empty((n_vox, plan.n_maps), dtype=torch.float64, device="cuda")
    for _ in range(2):
        plan.fit(y, d, l1, l2, out=est)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 3
    for _ in range(K):
        plan.fit(y, d, l1, l2, out=est)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    m = plan.m
    B = 4 * m + 24 + 4 * m * (plan.n_atoms - 1) + 4 * plan.n_maps
    r = {"model": model, "cfg": cfg, "m": m, "n_atoms": plan.n_atoms, "n_vox": n_vox, "ms": ms, "voxels_per_s": n_vox / ms * 1e3,
         "kernel_ms": plan.last_timing()["fit_kernel_ms"], "counters": plan.last_counters(), "bytes_per_voxel": B,
         "setup_s": time.time() - t0}
    print(json.dumps(r), flush=True)
    out[f"{model}_cfg{cfg}"] = r
    plan.close()
    del y, d, est
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_models.json", "w"), indent=1)
