#!/usr/bin/env python
"""Headline benchmark: NODDI voxels/s on the named synthetic volume (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cfg 2] [--nvox N]

A *step* is one pass of the hot path (`model.fit`: LUT index, binning, fused per-voxel fit) over one
whole synthetic volume of config cfg2 (128x128x64 = 1,048,576 voxels, 2-shell 90-dir + 10 b0 -> m=100,
145 atoms, ndirs=500), per GPU (weak scaling: every rank fits its own volume; voxels are independent,
so there is no data-path collective -- KERNELS are broadcast once and the maps gathered once, outside
the timed steps).

`value`  : voxels/s, all ranks, inputs already resident in HBM (fp32 DWI, fp64 directions).
`e2e`    : the same through the host-buffer C-ABI call (pinned host y/dirs -> H2D -> fit -> D2H maps).
`roofline`: fused fit kernel, algorithmic bytes/voxel (4m + 24 + 4 m n_rot + 4 n_maps) / its
           CUDA-event duration, against the measured HBM copy peak (MEASURED_PEAKS.json).
`pipeline` (N=1): the callers either side of the fit -- amx_preprocess / amx_dti_directions / amx_scatter_maps GB/s and the
           whole raw-volume -> maps flow (tools/bench_pipeline.py); an extra, not the headline.
`cpu_baseline` (N=1): the CPU oracle (oracle/) on a bounded sample of the same workload, all host cores.
`--impl reference`: the reference's own CPU path -- daducci/AMICO's Cython `NODDI.fit` compiled
           unmodified (oracle/_ref; its absent third-party spams-cython solvers bound to the restated
           ones) -- on a bounded sample; falls back to the C oracle when oracle/_ref is not loadable.
"""
from __future__ import annotations

import argparse
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


def workload(cfg, n_vox):
    from amico_b200 import synth
    model, dims = synth.CONFIGS[cfg]
    scheme = synth.make_scheme(cfg)
    n_b0, m = scheme.b0_count, scheme.nS
    return {
        "workload": f"cfg{cfg}: {model} {dims[0]}x{dims[1]}x{dims[2]} synthetic volume, {len(scheme.shells)}-shell "
                    f"{scheme.dwi_count}-dir + {n_b0} b0 (m={m}), ndirs=500, default grids, Rician SNR 30",
        "voxels_per_gpu": int(n_vox),
        "l2_policy": "inputs larger than L2 (no flush needed)" if n_vox * m * 4 > 130e6 else "small input: L2-resident",
    }


def ncu_traffic(n_vox):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fit kernels for one launch, from the committed ncu --set full
    capture of this same command (profiles/); None when the capture is for another problem size."""
    p = os.path.join(ROOT, "profiles", "ncu_full_r01_noddi_stage_kernels.json")
    try:
        with open(p) as f:
            ks = json.load(f)
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


def cpu_fit(P, n_sample, kind, threads):
    """One timed CPU fit of the first n_sample voxels.  Returns seconds."""
    from amico_b200 import synth
    Q = synth.Problem(P.cfg, P.model, P.scheme, P.lut_dirs, P.htable, P.KERNELS, P.params, P.y[:n_sample], P.DIRs[:n_sample])
    if kind == "reference":
        from oracle import ref_runner
        t = time.time()
        ref_runner.fit_problem(Q, nthreads=threads)
        return time.time() - t
    from oracle import oracle as orc
    t = time.time()
    orc.fit_problem(Q, nthreads=threads)
    return time.time() - t


def cpu_kind():
    try:
        from oracle import ref_runner
        if ref_runner.available():
            ref_runner._models()
            return "reference"
    except Exception:
        pass
    return "port"


def cpu_sample_size(cores, n_vox, m):
    # ~1500 NODDI voxels/s/core at m=100 (measured); aim at ~15 s; the reference overruns its y_est scratch for
    # chunks shorter than m (tests/golden/make_golden.py), so keep n/cores >= m
    n = int(min(n_vox, max(cores * max(m, 128), min(262144, cores * 1500 * 15))))
    return n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from amico_b200 import synth
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    scheme_m = synth.make_scheme(args.cfg).nS
    n_sample = cpu_sample_size(cores, args.nvox, scheme_m)
    P = synth.make_problem(args.cfg, n_vox=n_sample)
    from oracle import oracle as orc
    orc.load()
    for _ in range(args.warmup):
        cpu_fit(P, min(n_sample, cores * max(scheme_m, 128)), kind, cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_fit(P, n_sample, kind, cores)
    v = n_sample * args.steps / t
    sample = (f"{n_sample} voxels/step of the cfg{args.cfg} volume; " +
              ("daducci/AMICO NODDI.fit Cython compiled unmodified (oracle/_ref), ThreadPoolExecutor over all cores; the "
               "absent spams-cython nnls/lasso are the restated solvers of oracle/amico_oracle.c" if kind == "reference"
               else "C restatement oracle/amico_oracle.c, pthreads over all cores"))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload(args.cfg, args.nvox),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--nvox", type=int, default=0, help="voxels per GPU (default: the whole cfg volume)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the raw-volume -> maps flow (pre-processing, DTI, fit, scatter)")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = synth.CONFIGS[args.cfg][0]
    n_vox = args.nvox
    # rank 0 builds the protocol + KERNELS; one NCCL broadcast hands them to every GPU
    P = parallel.broadcast_problem(args.cfg, n_vox, rank, world, dev)
    from amico_b200 import models as amx_models
    mid = "FreeWater" if model.startswith("FreeWater") else model
    mdl = getattr(amx_models, mid)()
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

    # ---- end to end through the host-buffer C ABI (pinned host inputs, maps read back)
    e2e = None
    if not args.no_e2e:
        y_np = y_host.numpy()
        d_np = None if d_host is None else d_host.numpy()
        est_host = torch.empty((n_vox, plan.n_maps), dtype=torch.float64).pin_memory()
        est_np = est_host.numpy()
        for _ in range(2):
            plan.fit(y_np, d_np, l1, l2, out=est_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            plan.fit(y_np, d_np, l1, l2, out=est_np)
            launches_e2e = plan.last_counters()["launches"]
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = y_np.nbytes + (0 if d_np is None else d_np.nbytes)
        d2h = est_np.nbytes + (0 if d_np is None else d_np.nbytes)
        e2e = {"value": n_vox * world * args.steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(dt.item()) / args.steps}
        # sanity: host path and device path agree
        if not np.array_equal(est_np, est_dev.cpu().numpy()):
            e2e["note"] = "host-path maps differ from device-path maps"

    # ---- the final gather of the maps (outside the timed steps; checksum keeps it honest)
    maps = parallel.gather_maps(est_dev, rank, world)
    checksum = float(maps.sum().item()) if maps is not None else None

    fit_kernel_name = ("amx::k_noddi_stage<1|2|3> (NNLS / LARS / NNLS+maps stage kernels, timed as one span)"
                       if mid == "NODDI" else "amx::k_fit (fused per-voxel fit)")
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
            "dtype": "f64", "data": "synthetic", "config": dict(workload(args.cfg, n_vox), parallelism=f"voxel shards x{world}"),
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(n_vox), "kernel": fit_kernel_name, "kernel_ms_per_launch": 1e3 * k_s,
                         "bytes_per_voxel": B, "compulsory_bytes_per_voxel": B0,
                         "achieved_compulsory_GBs": B0 * n_vox / k_s / 1e9, "peak_source": peak_src},
            "fit": {"tiles": counters["tiles"], "warps_per_cta": counters["warps_per_cta"], "smem_bytes": counters["smem_bytes"],
                    "tma_staged": counters["tma_staged"], "grid": counters["grid"], "overflow_voxels": counters["overflow_voxels"]},
            "maps_checksum": checksum,
        }
        if mid == "NODDI":  # the three stage kernels run at their own widths (defaults of amx_api.cu; the counters above describe stage 3)
            line["fit"]["noddi_stage_warps"] = [int(os.environ.get(f"AMX_STAGE{k}_WARPS", d)) for k, d in ((1, 32), (2, 32), (3, 24))]
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            n_s = cpu_sample_size(cores, n_vox, m)
            from oracle import oracle as orc
            orc.load()
            cpu_fit(P, min(n_s, cores * max(m, 128)), "port", cores)
            dt = cpu_fit(P, n_s, "port", cores)
            line["cpu_baseline"] = {"value": n_s / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {n_s} voxels of the same volume, oracle/amico_oracle.c, {cores} pthreads"}
        if world == 1 and not args.no_pipeline and mid == "NODDI":
            # the callers either side of the fit: per-kernel GB/s and the whole raw-volume -> maps flow
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_pipeline
                del plan
                line["pipeline"] = bench_pipeline.measure(args.cfg, steps=max(2, min(args.steps, 5)), P=P)
            except Exception as e:  # an extra, never the headline
                line["pipeline"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
