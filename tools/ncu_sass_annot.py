"""Join the per-SASS-instruction counters of an ncu report with nvdisasm's line info.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > dis.txt
    python tools/ncu_sass_annot.py sass.csv dis.txt <kernel substring> <mangled section substring> <n_vox> [min_count_per_voxel]
Prints the kernel's instructions in program order with executions per voxel, stall samples and the source line, and
per-source-line totals split by opcode.
"""
import csv, re, sys
from collections import defaultdict, Counter

sass_csv, dis, ksub, msub, nvox = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
minc = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
rows = []; kern = None
for r in csv.reader(open(sass_csv)):
    if not r: continue
    if r[0] == "Kernel Name": kern = r[1]; continue
    if r[0] == "Address": continue
    if ksub in kern: rows.append((int(r[0], 16), r[1].strip(), int(r[4] or 0), int(r[5] or 0)))
seen = set(); uniq = []
for x in rows:
    if x[0] in seen: continue
    seen.add(x[0]); uniq.append(x)
base = uniq[0][0]
# nvdisasm
lines = {}; cur = None; on = False
for l in open(dis):
    if l.startswith(".text."):
        on = msub in l; continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3)); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: lines[int(m.group(1), 16)] = cur
tot = sum(x[3] for x in uniq)
print("# instructions", len(uniq), "executed per voxel %.1f" % (tot / nvox))
per = defaultdict(Counter); samp = Counter()
for a, txt, s, e in uniq:
    loc = lines.get(a - base)
    op = txt.split()[1] if txt.startswith("@") else txt.split()[0]
    key = "%s:%d" % (loc[0], loc[1]) if loc else "?"
    per[key][op.split(".")[0]] += e; samp[key] += s
    if e / nvox >= minc: print("%6x %8.1f %6d  %-28s %s" % (a - base, e / nvox, s, key, txt[:90]))
print("\n# per source line: executions per voxel, stall samples, opcode mix")
ts = sum(samp.values())
for key, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:70]:
    t = sum(c.values())
    print("%-26s %7.1f %5.2f%%  smp %5.2f%%  %s" % (key, t / nvox, 100 * t / tot, 100 * samp[key] / max(ts, 1), " ".join("%s:%.0f" % (k, v / nvox) for k, v in c.most_common(8))))
