"""A/B of the NODDI kernel organisations on the GPU box: parity vs the CPU oracle, agreement with each other, throughput.

    python tools/ab_noddi.py [n_parity] [n_speed] [cfg]

Variants are selected through the library's environment switches (read at every fit call).  Test tooling: uses
oracle/ as the checker.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amico_b200 import synth  # noqa: E402
from amico_b200.plan import Plan  # noqa: E402
from oracle import oracle as orc  # noqa: E402

VARIANTS = {
    "warp": {"AMX_NODDI_W32": "0"},
    "group": {"AMX_NODDI_W32": "1"},
}


def setenv(d):
    for k in ("AMX_NODDI_W32",):
        os.environ.pop(k, None)
    os.environ.update(d)


def relerr(g, r):
    return np.abs(g - r) / np.maximum(np.abs(r), 1e-3)


def main():
    n_par = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    n_spd = int(sys.argv[2]) if len(sys.argv) > 2 else 1048576
    cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    extra_env = [a for a in sys.argv[4:] if "=" in a]
    for kv in extra_env:
        k, v = kv.split("=", 1)
        os.environ[k] = v
    out = {}
    P = synth.make_problem(cfg, n_vox=n_par)
    l1, l2 = orc.DEFAULT_LAMBDAS["NODDI"]
    t0 = time.time()
    ref = orc.fit_problem(P, extra=True, return_debug=True, nthreads=os.cpu_count())
    print(f"oracle: {n_par} voxels in {time.time() - t0:.1f} s", flush=True)
    plan = Plan("NODDI", P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx)
    got = {}
    for name, env in VARIANTS.items():
        setenv(env)
        try:
            g = plan.fit(P.y, np.array(P.DIRs, dtype=np.float64), l1, l2, extra=True)
        except Exception as e:
            print(name, "FAILED:", repr(e), flush=True)
            continue
        got[name] = g
        rel = relerr(g["estimates"], ref["estimates"])
        ok = (rel <= 1e-4).all(1)
        s = {"pass_frac": float(ok.mean()), "n_fail": int((~ok).sum()), "p50": float(np.median(rel)),
             "p99": float(np.percentile(rel, 99)), "max": float(rel.max()),
             "support_equal_frac": float((g["support"] == ref["support"]).mean()) if "support" in g else None,
             "mod_maxabs": float(np.abs(g["estimates_mod"] - ref["estimates_mod"]).max()) if "estimates_mod" in g else None,
             "counters": plan.last_counters()}
        out[name] = s
        print(name, json.dumps(s), flush=True)
    if len(got) == 2:
        a, b = got["warp"]["estimates"], got["group"]["estimates"]
        rel = relerr(b, a)
        print("group vs warp: bit-equal voxels %.5f, within 1e-4: %.5f, max %.2e" % (
            float((a == b).all(1).mean()), float((rel <= 1e-4).all(1).mean()), float(rel.max())), flush=True)
        bad = np.nonzero(~(rel <= 1e-4).all(1))[0][:8]
        for i in bad:
            print("   vox", i, "oracle", ref["estimates"][i], "warp", a[i], "group", b[i], flush=True)
    plan.close()

    # ---- throughput, device-resident inputs
    if n_spd > 0:
        import torch
        P = synth.make_problem(cfg, n_vox=n_spd)
        plan = Plan("NODDI", P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx)
        dev = torch.device("cuda", 0)
        y = torch.from_numpy(P.y).to(dev)
        d = torch.from_numpy(np.ascontiguousarray(P.DIRs, dtype=np.float64)).to(dev)
        est = torch.empty((n_spd, plan.n_maps), dtype=torch.float64, device=dev)
        for name, env in VARIANTS.items():
            setenv(env)
            try:
                for _ in range(3):
                    plan.fit(y, d, l1, l2, out=est)
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for _ in range(5):
                    plan.fit(y, d, l1, l2, out=est)
                ev1.record()
                torch.cuda.synchronize()
                ms = ev0.elapsed_time(ev1) / 5
                print(f"{name}: {n_spd / ms * 1e3 / 1e6:.2f} M voxels/s, {ms:.2f} ms/step, fit kernels "
                      f"{plan.last_timing()['fit_kernel_ms']:.2f} ms, {plan.last_counters()}", flush=True)
                out[name + "_speed"] = {"voxels_per_s": n_spd / ms * 1e3, "ms": ms}
            except Exception as e:
                print(name, "FAILED:", repr(e), flush=True)
        plan.close()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ab_noddi.json", "w") as f:
        json.dump(out, f, indent=1, default=str)


if __name__ == "__main__":
    main()
