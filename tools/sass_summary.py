#!/usr/bin/env python
"""Per-kernel SASS instruction histogram + register / spill table of the shipped library (no GPU needed).

    python tools/sass_summary.py [lib.so] > profiles/sass_rNN_summary.json

* `cuobjdump -sass` of libamico_b200.so: for every kernel entry the number of SASS instructions and the counts of the
  mnemonics that identify the Blackwell-relevant paths -- DMMA (FP64 tensor-core `mma.sync.m8n8k4.f64`), UBLKCP (`cp.async.bulk`,
  the 1-D TMA copy), SYNCS (mbarrier), REDUX (warp-wide integer reductions of the arg-max / arg-min), SHFL, DFMA/DMUL/DADD,
  LDG/STG/LDS/STS, MUFU (rsqrt / rcp seeds), BAR/WARPSYNC.
* `cuobjdump -res-usage`: registers, shared memory, spill-free stack frame per entry.

(`nvcc -Xptxas -v` prints spill stores/loads at build time; `python -m amico_b200.build --force -v` shows them.  The stack frame
here is the upper bound of those spills plus local arrays.)
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "amico_b200", "libamico_b200.so")
WATCH = ["DMMA", "UBLKCP", "SYNCS", "REDUX", "SHFL", "DFMA", "DMUL", "DADD", "MUFU", "LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOM", "RED", "BAR",
         "WARPSYNC", "HMMA", "UTMALDG", "UTCHMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, cur = collections.OrderedDict(), None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", ln)
    if m and cur:
        op = m.group(1)
        kern[cur]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w):
                kern[cur][w] += 1
                break
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage, cur = {}, None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
    if m and cur:
        usage[cur] = {"registers": int(m.group(1)), "stack_bytes": int(m.group(2)), "static_shared_bytes": int(m.group(3)), "local_bytes": int(m.group(4))}
names = demangle(list(kern))
out = []
for k, c in kern.items():
    d = {"kernel": names.get(k, k), "mangled": k, "sass_instructions": c["_total"], "code_bytes": 16 * c["_total"]}
    d.update({w: c[w] for w in WATCH if c[w]})
    d.update(usage.get(k, {}))
    out.append(d)
want = ("k_noddi_stage1_lean<5, 1024, 16>", "k_noddi_stage2_lean<5, 1024>", "k_noddi_stage3_tpv<5, 5>", "k_noddi_stage<3, 5, float, 768>", "k_lasso_batched<2, 1, float, 1024>",
        "k_lasso_batched<1, 1, float, 1024>", "k_lasso_batched<3, 1, double, 1024>", "k_fit<2, 1, float>", "k_noddi_exact<5, float>", "k_preprocess",
        "k_dti", "k_gram<float>", "k_lut")
head = [d for d in out if any(w.replace(" ", "") in d["kernel"].replace("(int)", "").replace(" ", "") for w in want)]
print(json.dumps({"library": os.path.relpath(lib, ROOT), "command": "cuobjdump -sass / -res-usage (tools/sass_summary.py)",
                  "what_the_mnemonics_are": {"DMMA": "mma.sync.m8n8k4.f64 (FP64 tensor core)", "UBLKCP": "cp.async.bulk (1-D TMA copy)",
                                             "SYNCS": "mbarrier", "REDUX": "warp-wide integer reduction (arg-max / arg-min keys)"},
                  "kernels_on_the_default_paths": head, "all_kernels": out}, indent=1))
