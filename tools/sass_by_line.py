"""Static SASS instruction count per CUDA source line of one kernel section (nvdisasm -g -c output)."""
import re, sys
from collections import Counter
pat = sys.argv[2]
on=False; cur=('?',0); cnt=Counter()
for ln in open(sys.argv[1]):
    if ln.startswith('//---') and '.text.' in ln:
        on = pat in ln; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.search(r'/\*[0-9a-f]{4,5}\*/\s+\S', ln): cnt[cur]+=1
for (f,l),c in sorted(cnt.items(), key=lambda kv:-kv[1])[:int(sys.argv[3]) if len(sys.argv)>3 else 40]:
    print(f"{f}:{l}", c)
