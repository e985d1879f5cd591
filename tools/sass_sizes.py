import re, sys
pat = sys.argv[2] if len(sys.argv) > 2 else 'k_fitILi0ELi5EfEE'
on=False; cur=None; sizes={}; order=[]
for ln in open(sys.argv[1]):
    if ln.startswith('//---') and '.text.' in ln:
        on = pat in ln; continue
    if not on: continue
    m = re.match(r'\s*(\$?_Z[\w$]+|[\w$.]+):\s*$', ln)
    if m and not m.group(1).startswith('.L'):
        cur = m.group(1); sizes[cur]=0; order.append(cur); continue
    if re.search(r'/\*[0-9a-f]{4,5}\*/\s+\S', ln) and cur: sizes[cur]+=1
tot=0
for k in order:
    if sizes[k]: print(k.split('$')[-1][:70], sizes[k]); tot+=sizes[k]
print('total', tot, 'instr', tot*16/1024, 'KB')
