#!/usr/bin/env python
"""NODDI parity against the CPU oracle at scale (GPU box): python tools/parity_at_scale.py [n_vox] [cfg]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amico_b200 import synth  # noqa: E402
from amico_b200.plan import Plan  # noqa: E402
from oracle import oracle as orc  # noqa: E402

n_vox = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
P = synth.make_problem(cfg, n_vox=n_vox, seed=4242)
ref = orc.fit_problem(P, return_debug=True, nthreads=os.cpu_count())
with Plan("NODDI", P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx) as plan:
    got = plan.fit(P.y, np.array(P.DIRs), 0.5, 1e-3, debug=True)
rel = np.abs(got["estimates"] - ref["estimates"]) / np.maximum(np.abs(ref["estimates"]), 1e-3)
ok = (rel <= 1e-4).all(1)
print(json.dumps({"cfg": cfg, "voxels": n_vox, "within_1e-4": int(ok.sum()), "beyond": int((~ok).sum()), "pass_fraction": float(ok.mean()),
                  "support_equal": float((got["support"] == ref["support"]).mean()), "lut_equal": bool(np.array_equal(got["lut"], ref["lut"])),
                  "rel_err_p50": float(np.median(rel)), "rel_err_p99": float(np.percentile(rel, 99)), "rel_err_max": float(rel.max())}))
