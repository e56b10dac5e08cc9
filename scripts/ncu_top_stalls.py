#!/usr/bin/env python
"""Top stall locations (SASS) of one kernel from `ncu -i rep --page source --csv` output."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
h=[i for i,r in enumerate(rows) if 'Source' in r and '# Samples' in r][0]
hdr=rows[h]
i_src=hdr.index('Source'); i_s=hdr.index('# Samples'); i_ex=hdr.index('Instructions Executed')
stall=[i for i,c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
data=[]
for n,r in enumerate(rows[h+1:]):
    if len(r)<=i_s: continue
    try: s=int(r[i_s] or 0)
    except ValueError: continue
    top=sorted(((int(r[i] or 0),hdr[i]) for i in stall),reverse=True)[:2]
    data.append((s,r[i_src].strip(),int(r[i_ex] or 0),n,top))
tot=sum(d[0] for d in data)
print('total samples',tot,'instructions',len(data))
for s,src,ex,n,top in sorted(data,reverse=True)[:N]:
    print(f'{s:6d} {s/tot*100:5.1f}% line{n:5d} exec {ex:8d}  {src[:70]:70s} {top}')
