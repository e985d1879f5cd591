"""Summarise every kernel of an ncu report as JSON (the format kept under profiles/).

    python tools/ncu_to_json.py rep.ncu-rep "<command that was profiled>" [n_vox] > profiles/ncu_full_rNN_<what>.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
res = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    k = {key: d[key] for key in KEYS if key in d}
    k["unit"] = {key: u[key] for key in KEYS if key in u}
    st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(v), 2)
          for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and float(v) >= 0.1}
    k["stalls_warp_cycles_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1]))
    # scalar FP64 flop of the launch: thread-level DADD + DMUL + 2 DFMA (per-cycle sums over the sub-partitions x elapsed cycles);
    # the DMMA tensor flop are counted analytically by bench.py (2 m x padded atoms per voxel and A^T Y product)
    try:
        cyc = float(d["smsp__cycles_elapsed.avg"])
        per = {op: float(d[f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"]) for op in ("dadd", "dmul", "dfma")}
        k["fp64_flop_scalar"] = (per["dadd"] + per["dmul"] + 2.0 * per["dfma"]) * cyc
        k["smsp__cycles_elapsed.avg"] = cyc
    except (KeyError, ValueError):
        pass
    if len(sys.argv) > 3:
        k["n_vox"] = int(sys.argv[3])
    k["command"] = sys.argv[2] if len(sys.argv) > 2 else ""
    res.append(k)
print(json.dumps(res, indent=1))
