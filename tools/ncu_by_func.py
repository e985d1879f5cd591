"""Samples / executed instructions per device sub-function of one kernel.
usage: ncu_by_func.py sass.csv disasm.txt <kernel section pattern>"""
import csv, re, sys
sass_csv, disasm, pat = sys.argv[1:4]
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
iN, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
inst = [(int(r[iN] or 0), int(r[iE] or 0)) for r in rows[2:] if len(r) > iE]
on = False; cur = None; funcs = []
for ln in open(disasm):
    if ln.startswith('//---') and '.text.' in ln:
        on = pat in ln; continue
    if not on: continue
    m = re.match(r'\s*(\$?_Z[\w$]+|[\w$.]+):\s*$', ln)
    if m and not m.group(1).startswith('.L'):
        cur = m.group(1).split('$')[-1]; continue
    if re.search(r'/\*[0-9a-f]{4,5}\*/\s+\S', ln): funcs.append(cur)
assert len(funcs) == len(inst), (len(funcs), len(inst))
agg = {}
for f, (s, e) in zip(funcs, inst):
    a = agg.setdefault(f, [0, 0, 0]); a[0] += s; a[1] += e; a[2] += 1
ts = sum(a[0] for a in agg.values()); te = sum(a[1] for a in agg.values())
for f, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{f[:60]:60s} samples {100*a[0]/ts:6.2f}%  inst {100*a[1]/te:6.2f}%  ({a[1]/1e6:9.1f} M)  sass {a[2]}")
print("total inst", te)
