"""First-light / triage script for the GPU box: every model, small sizes, compared with the CPU oracle.

Usage (under gpurun):  python tools/gpu_first_light.py [n_vox]
Writes a JSON summary to gpurun_out/first_light.json.  Test tooling: uses oracle/ as the checker.
"""
import json
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amico_b200 import synth  # noqa: E402
from amico_b200.plan import Plan  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def compare(name, ref, got, tol=1e-4):
    r, g = ref["estimates"], got["estimates"]
    rel = np.abs(g - r) / np.maximum(np.abs(r), 1e-3)
    ok = (rel <= tol).all(1)
    out = {"pass_frac": float(ok.mean()), "n_fail": int((~ok).sum()), "bit_equal_frac": float((g == r).all(1).mean()),
           "p50": float(np.median(rel)), "p99": float(np.percentile(rel, 99)), "max": float(rel.max())}
    for k in ("rmse", "nrmse", "estimates_mod", "y_corrected"):
        if k in ref:
            d = np.abs(np.asarray(got[k]) - ref[k]).max()
            out["maxabs_" + k] = float(d)
    if "lut" in got and "lut" in ref:
        out["lut_equal"] = bool((got["lut"] == ref["lut"]).all())
    if "support" in got and "support" in ref:
        out["support_equal_frac"] = float((got["support"] == ref["support"]).mean())
    print(name, json.dumps(out), flush=True)
    if (~ok).any():
        bad = np.nonzero(~ok)[0][:5]
        for b in bad:
            print("   vox", b, "ref", r[b], "gpu", g[b], flush=True)
    return out


def main():
    n_vox = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    summary = {}
    cases = [(1, "FreeWater"), (1, "FreeWaterMouse"), (5, "CylinderZeppelinBall"), (4, "SANDI"), (2, "NODDI")]
    for cfg, model in cases:
        try:
            t0 = time.time()
            P = synth.make_problem(cfg, n_vox=n_vox, model=model)
            l1, l2 = orc.DEFAULT_LAMBDAS[model]
            ref = orc.fit_problem(P, rmse=True, nrmse=True, extra=True, return_debug=True, nthreads=os.cpu_count())
            t1 = time.time()
            mid = "FreeWater" if model.startswith("FreeWater") else model
            plan = Plan(mid, P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx)
            t2 = time.time()
            dirs = None if P.DIRs is None or model == "SANDI" else np.array(P.DIRs, dtype=np.float64)
            got = plan.fit(P.y, dirs, l1, l2, rmse=True, nrmse=True, extra=True, debug=True)
            t3 = time.time()
            s = compare(f"{model}(cfg{cfg}, {n_vox} vox)", ref, got)
            s.update(oracle_s=t1 - t0, plan_s=t2 - t1, fit_s=t3 - t2, timing=plan.last_timing(), counters=plan.last_counters())
            print("   ", s["timing"], s["counters"], flush=True)
            if dirs is not None:
                s["dirs_flipped_equal"] = bool((dirs == ref["dirs"]).all())
            # float64 input path must give the same answer
            got64 = plan.fit(P.y.astype(np.float64), None if dirs is None else np.array(P.DIRs, dtype=np.float64), l1, l2)
            s["f64_input_equal"] = bool((got64["estimates"] == got["estimates"]).all())
            summary[model] = s
            plan.close()
        except Exception as e:  # keep going: one broken model must not hide the others
            traceback.print_exc()
            summary[model] = {"error": repr(e)}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/first_light.json", "w") as f:
        json.dump(summary, f, indent=1, default=str)


if __name__ == "__main__":
    main()
