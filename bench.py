#!/usr/bin/env python
"""Headline benchmark: NODDI voxels/s on the named synthetic volume (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cfg 2] [--nvox N]

A *step* is one pass of the hot path (`model.fit`: LUT index, binning, per-voxel fit) over one whole synthetic volume
of config cfg2 (128x128x64 = 1,048,576 voxels, 2-shell 90-dir + 10 b0 -> m=100, 145 atoms, ndirs=500), per GPU (weak
scaling: every rank fits its own volume; voxels are independent, so there is no data-path collective -- KERNELS are
broadcast once and the maps gathered once, outside the timed steps).

`value`      : voxels/s, all ranks, inputs already resident in HBM (fp32 DWI, fp64 directions).
`e2e`        : the same through the host-buffer C-ABI call (pinned host y/dirs -> H2D -> fit -> D2H maps).
`e2e_plugin` : the same through the reference-facing plugin call `amico_b200.models.NODDI().fit(evaluation)` with
               `evaluation.y` float64 and PAGEABLE, exactly what `amico/core.py:451-452` hands a model.
`roofline`   : fit kernels, algorithmic bytes/voxel (4m + 24 + 4 m n_rot + 4 n_maps) / their CUDA-event
               duration, against the measured HBM copy peak (MEASURED_PEAKS.json).
`roofline_fp64`: the honest bound of these kernels -- FP64 flop and warp instructions per voxel (from the committed ncu
               capture) x voxels/s against the FP64 FMA / DMMA / issue peaks measured live (tools/fp64_peak.cu).
`configs`    (N=1): the other GPU configs of BASELINE.json -- cfg3 (NODDI 256x256x160, m=288), cfg4 (SANDI 128^3 with the
               fused directional average of the 192 raw volumes), cfg5 (CylinderZeppelinBall 200^3) -- each at its full
               size: voxels/s, roofline fraction, a parity sample against the oracle, the CPU baseline.
`sharded_cfg3` (N>1): ONE cfg3 volume strong-scaled over the ranks in voxel slabs: broadcast_ms, per-rank fit time,
               gather_ms and the sharded end-to-end throughput.
`pipeline`   (N=1): the callers either side of the fit (tools/bench_pipeline.py); an extra, not the headline.
`cpu_baseline` (N=1): the reference's CPU path on a bounded sample of the same workload, all host cores.
`--impl reference`: the reference's own CPU path -- daducci/AMICO's Cython `NODDI.fit` compiled unmodified (oracle/_ref;
               its absent third-party spams-cython solvers bound to the restated ones).  If oracle/_ref cannot be loaded the
               line says so (`cpu_baseline.kind: "port"`, `fallback_reason`) instead of degrading silently.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "NODDI voxels/sec (whole volume)"
UNIT = "voxels/s"
ROUND = "r02"


def workload(cfg, n_vox, world):
    """The `config` object -- identical in both arms."""
    from amico_b200 import synth
    model, dims = synth.CONFIGS[cfg]
    scheme = synth.make_scheme(cfg)
    n_b0, m = scheme.b0_count, scheme.nS
    return {
        "workload": f"cfg{cfg}: {model} {dims[0]}x{dims[1]}x{dims[2]} synthetic volume, {len(scheme.shells)}-shell "
                    f"{scheme.dwi_count}-dir + {n_b0} b0 (m={m}), ndirs=500, default grids, Rician SNR 30",
        "voxels_per_gpu": int(n_vox),
        "l2_policy": "inputs larger than L2 (no flush needed)" if n_vox * m * 4 > 130e6 else "small input: L2-resident",
        "parallelism": f"voxel shards x{world}",
    }


def ncu_summary():
    """Per-launch figures of the stage kernels from the committed ncu --set full capture of this command (profiles/)."""
    for name in (f"ncu_full_{ROUND}_noddi_stage_kernels.json", "ncu_full_r01_noddi_stage_kernels.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                with open(p) as f:
                    return json.load(f), name
            except Exception:
                pass
    return None, None


def ncu_traffic(n_vox):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fit kernels for one launch; None when the capture is for
    another problem size."""
    ks, _ = ncu_summary()
    try:
        if ks and int(ks[0].get("n_vox", 1048576)) == int(n_vox):
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            return float(sum(float(k[m]) * scale[k["unit"][m]] for k in ks for m in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
    except Exception:
        pass
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peaks(device):
    """FP64 FMA / DMMA TFLOP/s and warp-instruction issue rate measured live (tools/fp64_peak.cu)."""
    so = os.path.join(ROOT, "tools", "libfp64peak.so")
    if not os.path.exists(so):
        return None
    lib = C.CDLL(so)
    lib.fp64_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    out = {}
    for kind, key in ((0, "dfma_tflops"), (1, "dmma_tflops"), (2, "issue_tera_warp_inst_per_s")):
        v = C.c_double(0)
        rc = lib.fp64_peak(device, kind, C.byref(v))
        out[key] = v.value if rc == 0 else None
    return out


def bind_to_gpu_numa_node(gpu_index):
    """Multi-rank runs: pin this process (and the pinned host buffers it first-touches afterwards) to the CPUs NVML reports as
    local to its GPU, so that eight ranks do not stream their H2D traffic out of one NUMA node.  Best effort: returns the CPU
    list, or None when NVML / the affinity call is unavailable or the GPU reports every CPU (one node: nothing to gain)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < ncpu]
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if not cpus or len(cpus) >= len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arms
def cpu_fit(P, n_sample, kind, threads):
    """One timed CPU fit of the first n_sample voxels.  Returns (seconds, result dict)."""
    from amico_b200 import synth
    Q = synth.Problem(P.cfg, P.model, P.scheme, P.lut_dirs, P.htable, P.KERNELS, P.params, P.y[:n_sample],
                      None if P.DIRs is None else P.DIRs[:n_sample])
    if kind == "reference":
        from oracle import ref_runner
        t = time.time()
        r = ref_runner.fit_problem(Q, nthreads=threads)
        return time.time() - t, r
    from oracle import oracle as orc
    t = time.time()
    r = orc.fit_problem(Q, nthreads=threads)
    return time.time() - t, r


def cpu_kind():
    """("reference", None) when the reference's own Cython glue (oracle/_ref) loads; otherwise ("port", why)."""
    try:
        from oracle import ref_runner
        if not ref_runner.available() and os.path.isdir("/root/reference/amico"):
            from oracle import build_ref
            build_ref.build()
        if not ref_runner.available():
            d = os.path.join(ROOT, "oracle", "_ref", "amico")
            have = sorted(os.listdir(d)) if os.path.isdir(d) else None
            return "port", f"oracle/_ref has no built reference modules on this box (oracle/_ref/amico: {have})"
        ref_runner._models()
        return "reference", None
    except BaseException as e:  # SystemExit from the reference's ERROR() included
        return "port", f"{type(e).__name__}: {e}"


SAMPLE_TEXT = {
    "reference": "daducci/AMICO <Model>.fit Cython compiled unmodified (oracle/_ref), ThreadPoolExecutor over all cores; the "
                 "absent spams-cython nnls/lasso are the restated solvers of oracle/amico_oracle.c",
    "port": "C restatement oracle/amico_oracle.c, pthreads over all cores",
}


def cpu_sample_size(cores, n_vox, m, model="NODDI"):
    # ~1500 NODDI voxels/s/core at m=100 (measured); aim at ~15 s; the reference overruns its y_est scratch for
    # chunks shorter than m (tests/golden/make_golden.py), so keep n/cores >= m
    per_core = {"NODDI": 1500 * 100 / max(m, 1), "CylinderZeppelinBall": 4000, "SANDI": 60000, "FreeWater": 40000}.get(model, 1500)
    n = int(min(n_vox, max(cores * max(m, 128), min(262144, cores * per_core * 12))))
    return n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from amico_b200 import synth
    cores = os.cpu_count() or 1
    kind, why = cpu_kind()
    scheme_m = synth.make_scheme(args.cfg).nS
    n_sample = cpu_sample_size(cores, args.nvox, scheme_m)
    P = synth.make_problem(args.cfg, n_vox=n_sample)
    from oracle import oracle as orc
    orc.load()
    for _ in range(args.warmup):
        cpu_fit(P, min(n_sample, cores * max(scheme_m, 128)), kind, cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_fit(P, n_sample, kind, cores)[0]
    v = n_sample * args.steps / t
    sample = f"{n_sample} voxels/step of the cfg{args.cfg} volume; " + SAMPLE_TEXT[kind]
    base = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    if why:
        base["fallback_reason"] = why
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload(args.cfg, args.nvox, args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- helpers of our arm
def timed_steps(fn, steps, warmup):
    import torch
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


def parity_stats(got, ref):
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)
    ok = (rel <= 1e-4).all(axis=1)
    return {"voxels": int(len(ok)), "within_1e-4": int(ok.sum()), "pass_fraction": float(ok.mean()), "rel_err_p99": float(np.percentile(rel, 99)),
            "rel_err_max": float(rel.max()), "bit_equal": bool(np.array_equal(got, ref))}


def config_record(cfg, dev, steps, warmup, kind, cores, n_vox=None, cpu=True):
    """One of BASELINE.json's other GPU configs at its full size on this GPU: device-resident voxels/s, roofline fraction, a
    parity sample against the CPU arm on the first voxels, and the CPU baseline (bounded sample)."""
    import torch
    from amico_b200 import _lib as L
    from amico_b200 import models as amx_models, synth
    from amico_b200.plan import Plan
    model, dims = synth.CONFIGS[cfg]
    n_vox = int(n_vox or np.prod(dims))
    peak, _ = peaks()
    t0 = time.time()
    P = synth.make_problem(cfg, n_vox=8)  # protocol + KERNELS (voxels come from the GPU generator below)
    mdl = getattr(amx_models, model)()
    mdl.set_solver()
    l1, l2 = mdl.solver_params["lambda1"], mdl.solver_params["lambda2"]
    plan = Plan(model, P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx, device=dev.index)
    rec = {"cfg": cfg, "model": model, "voxels": n_vox, "m": plan.m, "n_atoms": plan.n_atoms}
    est = torch.empty((n_vox, plan.n_maps), dtype=torch.float64, device=dev)
    if model == "SANDI":
        # fused flow: raw 192-volume data -> amx_preprocess (b0 normalisation + directional average, SURVEY f-2) -> fit
        full = synth.make_scheme(cfg)
        raw = synth.make_sandi_raw_torch(full, n_vox, 20251017 + cfg, dev)
        shells = sorted(full.shells, key=lambda s: s["b"])
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        b0_idx, dwi_idx = i32(full.b0_idx), i32(full.dwi_idx)
        sh_idx = i32(np.concatenate([s["idx"] for s in shells]))
        sh_off = i32(np.concatenate([[0], np.cumsum([len(s["idx"]) for s in shells])]))
        m_out = 1 + len(shells)
        y = torch.empty((n_vox, m_out), dtype=torch.float32, device=dev)
        vox_idx = torch.empty(n_vox, dtype=torch.int32, device=dev)
        mb = torch.empty(n_vox, dtype=torch.float32, device=dev)
        a = L.PreArgs()
        a.space, a.device, a.dwi, a.n_total, a.nS = L.SPACE_DEVICE, dev.index, raw.data_ptr(), n_vox, full.nS
        a.b0_idx, a.b0_count, a.dwi_idx, a.dwi_count = b0_idx.ctypes.data, len(b0_idx), dwi_idx.ctypes.data, len(dwi_idx)
        a.shell_idx, a.shell_off, a.n_shells = sh_idx.ctypes.data, sh_off.ctypes.data, len(shells)
        a.flags, a.y, a.y_capacity, a.vox_idx, a.mean_b0s = L.PRE_NORMALIZE | L.PRE_DIR_AVG, y.data_ptr(), n_vox, vox_idx.data_ptr(), mb.data_ptr()
        a.stream = torch.cuda.current_stream(dev).cuda_stream
        kept, mo = C.c_int64(0), C.c_int(0)
        lib = L.load()

        def step():
            L.check(lib.amx_preprocess(C.byref(a), C.byref(kept), C.byref(mo)))
            plan.fit(y, None, l1, l2, out=est)

        B = 4 * full.nS + 24
        rec["what"] = "amx_preprocess (b0 normalisation + directional average of the 192 raw volumes) + amx_fit, device-resident"
        d = None
    else:
        y, d = synth.make_voxels_torch(model, P.KERNELS, P.htable, n_vox, 20251017 + cfg, dev)

        def step():
            plan.fit(y, d, l1, l2, out=est)

        B = 4 * plan.m + 24 + 4 * plan.m * (plan.n_atoms - 1) + 4 * plan.n_maps
        rec["what"] = "amx_fit, device-resident"
    rec["setup_s"] = round(time.time() - t0, 2)
    ms = timed_steps(step, steps, warmup)
    k_ms = plan.last_timing()["fit_kernel_ms"]
    cnt = plan.last_counters()
    rec.update({"ms_per_step": ms, "voxels_per_s": n_vox / ms * 1e3, "fit_kernel_ms": k_ms, "bytes_per_voxel": B,
                "roofline_frac": B * n_vox / (ms * 1e-3) / 1e9 / peak, "overflow_voxels": cnt["overflow_voxels"],
                "slow_path_voxels": cnt["slow_path_voxels"], "exact_path_voxels": cnt["exact_path_voxels"],
                "nan_values": int(torch.isnan(est).sum().item()), "gpu_launches_per_step": cnt["launches"] + (3 if model == "SANDI" else 0)})
    if cpu:
        # parity sample + CPU baseline on the first voxels of the same data (the CPU arm reads the fit's own inputs)
        n_s = cpu_sample_size(cores, n_vox, plan.m, model)
        y_s = y[:n_s].cpu().numpy()
        d_s = None if d is None else d[:n_s].cpu().numpy()
        Q = synth.Problem(cfg, model, P.scheme, P.lut_dirs, P.htable, P.KERNELS, P.params, y_s, d_s)
        dt, ref = cpu_fit(Q, n_s, kind, cores)
        rec["cpu_baseline"] = {"value": n_s / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": f"first {n_s} voxels; " + SAMPLE_TEXT[kind]}
        rec["parity_vs_cpu"] = parity_stats(est[:n_s].cpu().numpy(), np.asarray(ref["estimates"]))
    plan.close()
    return rec


def sharded_cfg3(rank, world, dev, steps):
    """ONE cfg3 volume (256x256x160, m=288) strong-scaled over the ranks in contiguous voxel slabs: rank 0 builds
    the KERNELS, one NCCL broadcast ships them (timed), every rank fits its slab (timed per rank), one NCCL gather returns the
    float32 maps to rank 0 (timed).  Each rank generates only its own slab of the volume."""
    import torch
    import torch.distributed as dist
    from amico_b200 import models as amx_models, parallel, synth
    from amico_b200.plan import Plan
    cfg = 3
    n_total = int(np.prod(synth.CONFIGS[cfg][1]))
    i0, i1 = parallel.shard_bounds(n_total, world, rank)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    P = parallel.broadcast_problem(cfg, 8, rank, world, dev)
    torch.cuda.synchronize(); dist.barrier()
    t_bcast = time.perf_counter() - t0  # includes rank 0's kernel synthesis; the NCCL part alone is timed below
    kbytes = sum(np.asarray(v).nbytes for k, v in P.KERNELS.items() if k != "model")
    # the NCCL broadcast alone: the tables as device tensors, CUDA events around the collectives
    dts = [torch.from_numpy(np.ascontiguousarray(v).view(np.uint8).reshape(-1)).to(dev) for k, v in sorted(P.KERNELS.items()) if k != "model"]
    for _ in range(2):
        for t in dts:
            dist.broadcast(t, src=0)
    torch.cuda.synchronize(); dist.barrier()
    eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eb0.record()
    for t in dts:
        dist.broadcast(t, src=0)
    eb1.record()
    torch.cuda.synchronize()
    tb = torch.tensor([eb0.elapsed_time(eb1)], dtype=torch.float64, device=dev)
    dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    t_bcast_nccl = float(tb.item()) / 1e3
    del dts
    mdl = amx_models.NODDI()
    mdl.set_solver()
    l1, l2 = mdl.solver_params["lambda1"], mdl.solver_params["lambda2"]
    plan = Plan("NODDI", P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx, device=dev.index)
    y, d = synth.make_voxels_torch("NODDI", P.KERNELS, P.htable, i1 - i0, 20251017 + cfg + 1000 * rank, dev)
    est = torch.empty((i1 - i0, 3), dtype=torch.float64, device=dev)
    for _ in range(2):
        plan.fit(y, d, l1, l2, out=est)
    gather_ms, fit_ms, tot_ms = [], [], []
    for _ in range(steps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        plan.fit(y, d, l1, l2, out=est)
        e1.record()
        maps = parallel.gather_maps(est.to(torch.float32), rank, world)
        e2.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2), e0.elapsed_time(e2)], dtype=torch.float64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        fit_ms.append([float(a[0]) for a in allt]); gather_ms.append(max(float(a[1]) for a in allt)); tot_ms.append(max(float(a[2]) for a in allt))
    # diagnostics: non-finite values in this rank's own maps (before the gather) next to those in the gathered volume
    nf = torch.tensor([float((~torch.isfinite(est)).sum().item())], dtype=torch.float64, device=dev)
    nf_all = [torch.empty_like(nf) for _ in range(world)]
    dist.all_gather(nf_all, nf)
    plan.close()
    if rank != 0:
        return None
    tot = float(np.mean(tot_ms))
    return {"workload": "cfg3: ONE NODDI 256x256x160 volume (10,485,760 voxels, m=288) in contiguous voxel slabs, strong scaling",
            "voxels": n_total, "ranks": world, "kernels_bytes": int(kbytes), "setup_broadcast_ms_incl_synthesis": 1e3 * t_bcast,
            "broadcast_ms": 1e3 * t_bcast_nccl, "fit_ms_per_rank": [float(x) for x in np.mean(np.array(fit_ms), axis=0)],
            "gather_ms": float(np.mean(gather_ms)), "gather_bytes": int(n_total * 3 * 4), "ms_per_volume": tot,
            "voxels_per_s": n_total / tot * 1e3, "checksum": float(torch.nan_to_num(maps).sum().item()), "nan_values": int(torch.isnan(maps).sum().item()),
            "nonfinite_per_rank_before_gather": [int(x.item()) for x in nf_all],
            "what": "per step: barrier, fit of the rank's slab (device-resident), NCCL gather of the float32 maps on rank 0; max over ranks"}


# --------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--nvox", type=int, default=0, help="voxels per GPU (default: the whole cfg volume)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the raw-volume -> maps flow (pre-processing, DTI, fit, scatter)")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3 / cfg4 / cfg5 sub-records and the sharded cfg3 run")
    args = ap.parse_args()
    from amico_b200 import synth
    if not args.nvox:
        args.nvox = int(np.prod(synth.CONFIGS[args.cfg][1]))
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from amico_b200.plan import Plan
    from amico_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: amico_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 and os.environ.get("AMX_BENCH_BIND", "1") != "0" else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = synth.CONFIGS[args.cfg][0]
    n_vox = args.nvox
    # rank 0 builds the protocol + KERNELS; one NCCL broadcast hands them to every GPU
    P = parallel.broadcast_problem(args.cfg, n_vox, rank, world, dev)
    from amico_b200 import models as amx_models
    mid = "FreeWater" if model.startswith("FreeWater") else model
    mdl = getattr(amx_models, mid)()
    mdl.scheme = P.scheme
    mdl.set_solver()  # the reference's default lambdas (amico/models.pyx:721, :1077, :439, :1405)
    l1, l2 = mdl.solver_params["lambda1"], mdl.solver_params["lambda2"]
    plan = Plan(mid, P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx, device=local)

    y_host = torch.from_numpy(P.y).pin_memory()
    y_dev = y_host.to(dev, non_blocking=True)
    d_host = d_dev = None
    if P.DIRs is not None and mid != "SANDI":
        d_host = torch.from_numpy(np.ascontiguousarray(P.DIRs, dtype=np.float64)).pin_memory()
        d_dev = d_host.to(dev, non_blocking=True)
    est_dev = torch.empty((n_vox, plan.n_maps), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        plan.fit(y_dev, d_dev, l1, l2, out=est_dev)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, launches = 0.0, 0
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
        kernel_ms += plan.last_timing()["fit_kernel_ms"]
        launches += plan.last_counters()["launches"]
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    counters = plan.last_counters()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n_vox * world * args.steps / (ms_max / 1e3)

    def wall_steps(fn, reps):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item())

    # ---- end to end through the host-buffer C ABI (pinned host inputs, maps read back)
    e2e = e2e_plugin = None
    if not args.no_e2e:
        y_np = y_host.numpy()
        d_np = None if d_host is None else d_host.numpy()
        est_host = torch.empty((n_vox, plan.n_maps), dtype=torch.float64).pin_memory()
        est_np = est_host.numpy()
        for _ in range(2):
            plan.fit(y_np, d_np, l1, l2, out=est_np)
        dt = wall_steps(lambda: plan.fit(y_np, d_np, l1, l2, out=est_np), args.steps)
        h2d = y_np.nbytes + (0 if d_np is None else d_np.nbytes)
        d2h = est_np.nbytes + (0 if d_np is None else d_np.nbytes)
        e2e = {"value": n_vox * world * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * dt / args.steps,
               "what": "Plan.fit on pinned float32 host buffers (amx_fit, AMX_SPACE_HOST): chunked H2D | fit | D2H pipeline"}
        # sanity: host path and device path agree
        if not np.array_equal(est_np, est_dev.cpu().numpy()):
            e2e["note"] = "host-path maps differ from device-path maps"

        # ---- the reference-facing plugin call: model.fit(evaluation), evaluation.y float64 and pageable (core.py:451-452)
        class _Evaluation:
            def __init__(self):
                self.y = P.y.astype(np.float64)  # pageable float64, float32-valued like the reference's
                self.DIRs = None if P.DIRs is None else np.array(P.DIRs, dtype=np.float64)
                self.htable, self.KERNELS, self.nthreads = P.htable, P.KERNELS, -1

            def get_config(self, k):
                return None

        ev = _Evaluation()
        mdl.device = local
        res = None
        for _ in range(2):
            res = mdl.fit(ev)

        def plug():
            nonlocal res
            res = mdl.fit(ev)

        reps = max(2, min(args.steps, 5))
        dt = wall_steps(plug, reps)
        e2e_plugin = {"value": n_vox * world * reps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / reps,
                      "h2d_bytes_per_step": int(ev.y.nbytes + (0 if ev.DIRs is None else ev.DIRs.nbytes)),
                      "d2h_bytes_per_step": int(res["estimates"].nbytes + (0 if ev.DIRs is None else ev.DIRs.nbytes)),
                      "host_bytes_read_per_step": int(ev.y.nbytes),
                      "what": "amico_b200.models.%s().fit(evaluation): evaluation.y float64 pageable (amico/core.py:451-452), "
                              "returns the reference's dict of float64 host arrays" % mid,
                      "maps_equal_device_path": bool(np.array_equal(res["estimates"], est_dev.cpu().numpy()))}

    # ---- the final gather of the maps (outside the timed steps; checksum keeps it honest)
    maps = parallel.gather_maps(est_dev, rank, world)
    checksum = float(maps.sum().item()) if maps is not None else None

    fit_kernel_name = ("amx::k_noddi_stage1_lean | k_noddi_stage2_lean | k_noddi_stage3_tpv (NNLS / LARS / NNLS+maps stage kernels, timed as one span)"
                       if mid == "NODDI" else "amx::k_fit (fused per-voxel fit)")
    line = None
    plan_n_atoms = plan.n_atoms
    if rank == 0:
        m, n_maps = plan.m, plan.n_maps
        n_rot = plan.n_atoms - 1
        B = 4 * m + 24 + 4 * m * n_rot + 4 * n_maps
        B0 = 4 * m + 24 + 4 * n_maps
        peak, peak_src = peaks()
        k_s = kernel_ms / args.steps / 1e3
        achieved = B * n_vox / k_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload(args.cfg, n_vox, world),
            "clocks": clocks, "gpu_launches": int(launches),
            "host_cpu_binding": ("GPU-local CPUs %d-%d (%d)" % (numa[0], numa[-1], len(numa))) if numa else "none (single rank, one NUMA node, or NVML unavailable)",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(n_vox), "kernel": fit_kernel_name, "kernel_ms_per_launch": 1e3 * k_s,
                         "bytes_per_voxel": B, "compulsory_bytes_per_voxel": B0,
                         "achieved_compulsory_GBs": B0 * n_vox / k_s / 1e9, "peak_source": peak_src},
            "fit": {"tiles": counters["tiles"], "warps_per_cta": counters["warps_per_cta"], "smem_bytes": counters["smem_bytes"],
                    "tma_staged": counters["tma_staged"], "grid": counters["grid"], "overflow_voxels": counters["overflow_voxels"],
                    "slow_path_voxels": counters["slow_path_voxels"]},
            "maps_checksum": checksum,
        }
        if e2e:
            line["e2e"] = e2e
        if e2e_plugin:
            line["e2e_plugin"] = e2e_plugin
        # ---- FP64 / issue roofline: per-voxel counts from the committed ncu capture, peaks measured live
        pk = fp64_peaks(local)
        ks, ks_name = ncu_summary()
        if pk:
            r64 = {"peaks": pk, "peak_source": "tools/fp64_peak.cu run in this process"}
            try:
                nv = int(ks[0].get("n_vox", 1048576))
                # scalar FP64 flop from the ncu thread-instruction counters + the DMMA flop of the three A^T Y micro-GEMMs
                # (2 x rows x 32 NPL padded atoms per voxel: stage 1 over m rows, stage 2 over the DWI rows)
                # (stage 3 reads the c1 stage 1 stored, so two GEMMs per voxel: m rows + the DWI rows)
                dmma = 2.0 * (m + P.scheme.dwi_count) * 32 * ((plan_n_atoms + 31) // 32)
                flops = sum(float(k["fp64_flop_scalar"]) for k in ks) / nv + dmma
                winst = sum(float(k["smsp__inst_executed.sum"]) for k in ks) / nv
                vps = n_vox / k_s
                r64.update({"flops_per_voxel": flops, "dmma_flops_per_voxel": dmma, "warp_inst_per_voxel": winst, "counts_from": "profiles/" + ks_name,
                            "achieved_tflops": flops * vps / 1e12, "peak_tflops": pk["dfma_tflops"],
                            "frac": flops * vps / 1e12 / pk["dfma_tflops"] if pk["dfma_tflops"] else None,
                            "issue_utilisation": winst * vps / 1e12 / pk["issue_tera_warp_inst_per_s"] if pk["issue_tera_warp_inst_per_s"] else None})
            except Exception as e:
                r64["counts_error"] = repr(e)
            line["roofline_fp64"] = r64
    cores = os.cpu_count() or 1
    if rank == 0 and world == 1 and not args.no_cpu:
        kind, why = cpu_kind()
        n_s = cpu_sample_size(cores, n_vox, plan.m)
        from oracle import oracle as orc
        orc.load()
        cpu_fit(P, min(n_s, cores * max(plan.m, 128)), kind, cores)
        dt, _ = cpu_fit(P, n_s, kind, cores)
        line["cpu_baseline"] = {"value": n_s / dt, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"first {n_s} voxels of the same volume; " + SAMPLE_TEXT[kind]}
        if why:
            line["cpu_baseline"]["fallback_reason"] = why
    if rank == 0 and world == 1 and not args.no_pipeline and mid == "NODDI":
        # the callers either side of the fit: per-kernel GB/s and the whole raw-volume -> maps flow
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_pipeline
            line["pipeline"] = bench_pipeline.measure(args.cfg, steps=max(2, min(args.steps, 5)), P=P)
        except Exception as e:  # an extra, never the headline
            line["pipeline"] = {"error": repr(e)}
    del y_dev, d_dev, est_dev
    plan.close()
    torch.cuda.empty_cache()
    if not args.no_configs and args.cfg == 2:
        if world == 1:
            kind = "port" if args.no_cpu else cpu_kind()[0]
            recs = {}
            for c in (3, 4, 5):
                try:
                    recs[f"cfg{c}"] = config_record(c, dev, steps=max(2, min(args.steps, 3)), warmup=2, kind=kind, cores=cores, cpu=not args.no_cpu)
                except Exception as e:
                    recs[f"cfg{c}"] = {"error": repr(e)}
                torch.cuda.empty_cache()
            line["configs"] = recs
        else:
            try:
                rec = sharded_cfg3(rank, world, dev, steps=max(2, min(args.steps, 3)))
            except Exception as e:
                rec = {"error": repr(e)}
            if rank == 0:
                line["sharded_cfg3"] = rec
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
