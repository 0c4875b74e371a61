#!/bin/bash
# one rank's share of the 8-GPU run on one GPU (no NCCL): where does the step time go beyond the profiled categories?
timeout 300 python bench.py --rows 125000 --steps 20 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/q_shard.json 2> gpurun_out/q_shard.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_shard.json').read().strip().splitlines()[-1])
print('rows 125000: ms/step',round(d['ms_per_step'],3),'breakdown',{k:round(x,3) for k,x in d['breakdown_ms_per_step'].items()}, 'launches/step', d['gpu_launches']/d['steps'], 'e2e', d['e2e'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/shard_launches.csv python bench.py --rows 125000 --steps 1 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > /dev/null 2>&1
python - <<'PY'
import csv,re
rows=list(csv.reader(open('gpurun_out/shard_launches.csv')))
hdr=None; data=[]
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(dict(zip(hdr,r)))
names=[re.sub(r'\(.*','',d['Kernel Name']).replace('ggp::','')[:44] for d in data]
# the timed evaluation = the 4th k_build_kzz (3 warm-ups + 1 timed)... take the one before the probes: find indices
idx=[i for i,n in enumerate(names) if n.startswith('k_build_kzz')]
s=idx[3] if len(idx)>3 else idx[-1]
e=idx[4] if len(idx)>4 else len(data)
tot=0; agg={}
for i in range(s-3,min(e,s+140)):
    t=float(data[i]['Metric Value'].replace(',',''))/1000
    if 'probe' in names[i]: break
    tot+=t; a=agg.setdefault(names[i],[0,0.0]); a[0]+=1; a[1]+=t
print('one evaluation under ncu: %.1f us in %d kernel kinds' % (tot, len(agg)))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:30]: print('%-46s %3d %8.1f us' % (k, v[0], v[1]))
PY
