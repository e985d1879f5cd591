"""Aggregate an ncu SASS-level source page by CUDA source line.

    ncu -i rep.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > disasm.txt
    python tools/ncu_by_line.py sass.csv disasm.txt <mangled kernel name> [top]
"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, disasm, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
iA, iS, iN, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
inst = [(r[iS].strip(), int(r[iN] or 0), int(r[iE] or 0)) for r in rows[2:] if len(r) > iE]
# line info per instruction, in order
lines = []
on = False
cur = ("?", 0)
stack = []
for ln in open(disasm):
    if ln.startswith("//---") and ".text." in ln:
        on = kname in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        inl = m.group(3)
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in inl)
        continue
    if re.search(r"/\*[0-9a-f]{4,5}\*/\s+\S", ln):
        lines.append(cur)
print(len(inst), "profiled instructions;", len(lines), "disassembled")
n = min(len(inst), len(lines))
agg = defaultdict(lambda: [0, 0, 0])
for (src, smp, ex), loc in zip(inst[:n], lines[:n]):
    a = agg[loc[:2]]
    a[0] += smp; a[1] += ex; a[2] += 1
tot_s = sum(a[0] for a in agg.values()); tot_e = sum(a[1] for a in agg.values())
print("total samples", tot_s, "total instr executed", tot_e)
print("%-28s %8s %6s %14s %6s %6s" % ("file:line", "samples", "%", "inst_exec", "%", "sass"))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-28s %8d %6.2f %14d %6.2f %6d" % (f"{loc[0]}:{loc[1]}", a[0], 100 * a[0] / max(tot_s, 1), a[1], 100 * a[1] / max(tot_e, 1), a[2]))

# coarse grouping by file + line ranges given as extra args: name:file:lo-hi
groups = [g.split(":") for g in sys.argv[5:]]
if groups:
    print()
    gagg = defaultdict(lambda: [0, 0])
    for loc, a in agg.items():
        name = "other:" + loc[0]
        for gname, gfile, rng in groups:
            lo, hi = map(int, rng.split("-"))
            if loc[0] == gfile and lo <= loc[1] <= hi:
                name = gname
                break
        gagg[name][0] += a[0]; gagg[name][1] += a[1]
    for name, a in sorted(gagg.items(), key=lambda kv: -kv[1][0]):
        print("%-36s samples %6.2f%%  inst %6.2f%%" % (name, 100 * a[0] / tot_s, 100 * a[1] / tot_e))
