"""Per-source-line totals from `ncu -i rep --page source --csv --print-source cuda,sass` (needs -lineinfo + --import-source on).

    python tools/ncu_src_lines.py src.csv [top] [kernel-substring]
Prints, per kernel, the lines with the most executed warp instructions and their stall samples.
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
kern = None
fpath = None
hdr = None
agg = {}
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kern = r[1]
        agg.setdefault(kern, defaultdict(lambda: [0, 0, ""]))
        continue
    if r[0] == "Line No":
        hdr = r
        iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or kern is None or r[0] == "":
        continue
    try:
        a = agg[kern][(fpath, int(r[0]))]
        a[0] += int(r[iS] or 0)
        a[1] += int(r[iE] or 0)
        a[2] = r[1].strip()[:90]
    except (ValueError, IndexError):
        pass
for k, d in agg.items():
    if want not in k:
        continue
    ts = sum(a[0] for a in d.values())
    te = sum(a[1] for a in d.values())
    print(f"== {k}: samples {ts}, warp instructions {te}")
    for loc, a in sorted(d.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-24s inst %6.2f%%  samples %6.2f%%  %s" % (f"{loc[0]}:{loc[1]}", 100 * a[1] / max(te, 1), 100 * a[0] / max(ts, 1), a[2]))
