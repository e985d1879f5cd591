#!/usr/bin/env python
"""Throughput of the callers either side of the fit and of the whole
raw-volume -> maps flow (`amico_b200.evaluation.Evaluation`), on one GPU.

    python tools/bench_pipeline.py [--cfg 2] [--steps 5] [--json]

Per kernel: CUDA-event time of the device-space C-ABI call on torch's current stream, ALGORITHMIC bytes (what the pass must
read and write once) / time against the measured HBM copy peak.  Flow: pinned host volume -> H2D -> amx_preprocess ->
amx_dti_directions -> amx_fit -> amx_scatter_maps -> D2H of the float32 map volumes, wall clock.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(cfg=2, steps=5, warmup=3, n_vox=None, P=None):
    import torch
    from amico_b200 import _lib as L
    from amico_b200 import synth
    from amico_b200.evaluation import Evaluation, MIN_POSITIVE_SIGNAL

    lib = L.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    model, dims = synth.CONFIGS[cfg]
    if P is None:
        P = synth.make_problem(cfg, n_vox=n_vox or int(np.prod(dims)))
    n_vox = len(P.y)
    sch = P.scheme
    nS = sch.nS
    # raw volume: the problem's normalised signal times a per-voxel S0 (float32), all voxels in the mask
    rng = np.random.default_rng(5)
    s0 = rng.uniform(400.0, 1600.0, (n_vox, 1)).astype(np.float32)
    raw = torch.from_numpy((P.y * s0).astype(np.float32)).pin_memory()
    vol4d = raw.numpy().reshape(n_vox, 1, 1, nS)
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            peak = float(json.load(f)["hbm_gbs"])

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    out = {"cfg": cfg, "voxels": n_vox, "nS": nS, "hbm_peak_GBs": peak}
    # ---- amx_preprocess (device space)
    d_raw = raw.to(dev)
    y = torch.empty((n_vox, nS), dtype=torch.float32, device=dev)
    vidx = torch.empty(n_vox, dtype=torch.int32, device=dev)
    mb = torch.empty(n_vox, dtype=torch.float32, device=dev)
    b0, dw = np.ascontiguousarray(sch.b0_idx, np.int32), np.ascontiguousarray(sch.dwi_idx, np.int32)
    a = L.PreArgs()
    a.space, a.device, a.dwi, a.n_total, a.nS = L.SPACE_DEVICE, dev.index, d_raw.data_ptr(), n_vox, nS
    a.b0_idx, a.b0_count, a.dwi_idx, a.dwi_count = b0.ctypes.data, len(b0), dw.ctypes.data, len(dw)
    a.flags, a.y, a.y_capacity, a.vox_idx, a.mean_b0s = L.PRE_NORMALIZE, y.data_ptr(), n_vox, vidx.data_ptr(), mb.data_ptr()
    a.stream = torch.cuda.current_stream().cuda_stream
    kept, mo = C.c_int64(0), C.c_int(0)

    def pre():
        L.check(lib.amx_preprocess(C.byref(a), C.byref(kept), C.byref(mo)))

    ms = timed(pre)
    bytes_pre = n_vox * (2 * nS * 4 + 8)
    out["preprocess"] = {"ms": ms, "bytes": bytes_pre, "GBs": bytes_pre / ms / 1e6, "frac_of_hbm_peak": bytes_pre / ms / 1e6 / peak,
                         "what": "amx_preprocess call (mask count + scan + k_preprocess + 8-byte read-back), normalise only, mask = all"}
    # ---- amx_dti_directions
    from amico_b200.evaluation import dti_design_matrix
    W = np.ascontiguousarray(np.linalg.pinv(dti_design_matrix(sch.b, sch.raw[:, :3]))[:6])
    dirs = torch.empty((n_vox, 3), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def dti():
        L.check(lib.amx_dti_directions(dev.index, L.SPACE_DEVICE, y.data_ptr(), L.F32, n_vox, nS, W.ctypes.data,
                                       MIN_POSITIVE_SIGNAL, dirs.data_ptr(), st))

    ms = timed(dti)
    bytes_dti = n_vox * (nS * 4 + 24)
    out["dti"] = {"ms": ms, "bytes": bytes_dti, "GBs": bytes_dti / ms / 1e6, "frac_of_hbm_peak": bytes_dti / ms / 1e6 / peak,
                  "fp64_log_per_voxel": nS}
    # ---- amx_scatter_maps
    est = torch.rand((n_vox, 3), dtype=torch.float64, device=dev)
    volm = torch.empty((n_vox, 3), dtype=torch.float32, device=dev)

    def sca():
        L.check(lib.amx_scatter_maps(dev.index, L.SPACE_DEVICE, est.data_ptr(), n_vox, 3, vidx.data_ptr(), volm.data_ptr(), n_vox, st))

    ms = timed(sca)
    bytes_sc = n_vox * (24 + 4 + 12 + 12)
    out["scatter"] = {"ms": ms, "bytes": bytes_sc, "GBs": bytes_sc / ms / 1e6, "frac_of_hbm_peak": bytes_sc / ms / 1e6 / peak}
    del d_raw, y, vidx, mb, dirs, est, volm
    # ---- whole flow, host volume in, host maps out
    ae = Evaluation(device=dev.index)
    ae.set_model("FreeWater" if model.startswith("FreeWater") else model)

    def flow():
        ae.load_data(vol4d, sch)
        ae.load_kernels(P.KERNELS, P.htable)
        return ae.fit()

    for _ in range(warmup):
        flow()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = flow()
    dt = (time.perf_counter() - t0) / steps
    out["flow"] = {"voxels_per_s": n_vox / dt, "ms": 1e3 * dt, "h2d_bytes": int(raw.numel() * 4),
                   "d2h_bytes": int(sum(v.nbytes for v in res.values())),
                   "dirs_ms": 1e3 * ae.get_config("dirs_precomputing_time"), "fit_ms": 1e3 * ae.get_config("fit_time"),
                   "what": "Evaluation.load_data + fit: pageable->device upload of the raw volume, preprocess, DTI directions, "
                           "model fit, scatter, download of MAPs + DIRs"}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--nvox", type=int, default=0)
    args = ap.parse_args()
    print(json.dumps(measure(args.cfg, args.steps, n_vox=args.nvox or None)))
