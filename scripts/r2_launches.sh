#!/bin/bash
# launch list of ONE evaluation (the third of three) at a reduced row count: who is who in the m x m section
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_small.csv python scripts/prof_one_eval_i8.py 32768 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_small.csv')) if len(r)>5 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section Name, Metric Name, Metric Unit, Metric Value
n=len(rows)//3
last=rows[2*n:]
agg=collections.OrderedDict()
for r in last:
    name=r[4].split('(')[0][:60]; v=float(r[-1].replace(',',''))
    u=r[-2]
    if u=='ns': v/=1000
    elif u=='ms': v*=1000
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print('launches',len(last),'total us',round(tot,1))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f'{a[0]:4d} x {a[1]/a[0]:9.1f} us = {a[1]:9.1f} us  {k}')
PY
