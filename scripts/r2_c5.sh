#!/bin/bash
true
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c5_launches.csv python scripts/c5_eval.py ${C5_CHAINS:-8} 2 > gpurun_out/c5_ncu.log 2>&1
tail -3 gpurun_out/c5_ncu.log
python - <<'PY'
import csv,re
rows=list(csv.reader(open('gpurun_out/c5_launches.csv')))
hdr=None; data=[]
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(dict(zip(hdr,r)))
names=[re.sub(r'\(.*','',d['Kernel Name']).replace('ggp::','')[:40] for d in data]
half=len(data)//2
agg={}
for i in range(half,len(data)):
    t=float(data[i]['Metric Value'].replace(',',''))/1000
    a=agg.setdefault(names[i]+' '+data[i]['Grid Size'],[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(v[1] for v in agg.values())
print('second evaluation: %d launches, %.1f ms' % (len(data)-half, tot/1000))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print('%-60s %4d %9.1f us %5.1f%%' % (k, v[0], v[1], 100*v[1]/tot))
PY
