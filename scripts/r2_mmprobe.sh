#!/bin/bash
python scripts/mm_probe.py 1024 2>&1 | tail -7
GGP_MM64_MAX_TILES=0 GGP_CHOL_LOOKAHEAD=0 python scripts/mm_probe.py 1024 2>&1 | tail -6
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mm_probe_launches.csv python scripts/mm_probe.py 1024 1 > /dev/null 2>&1
python - <<'PY'
import csv,re
rows=list(csv.reader(open('gpurun_out/mm_probe_launches.csv')))
hdr=None; data=[]
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(dict(zip(hdr,r)))
names=[re.sub(r'\(.*','',d['Kernel Name']).replace('ggp::','') for d in data]
idx=[i for i,n in enumerate(names) if n.startswith('k_potf2_trti2')]
# last chol: from last standalone potf2 up to the final transpose
s=idx[-1]
tot=0; agg={}
for i in range(s,len(data)):
    d=data[i]; t=float(d['Metric Value'].replace(',',''))/1000
    if names[i].startswith('void at::') : continue
    agg.setdefault(names[i][:30],[]).append(round(t,1))
for k,v in agg.items(): print(k, len(v), round(sum(v),1), v if len(v)<=20 else v[:20])
PY
