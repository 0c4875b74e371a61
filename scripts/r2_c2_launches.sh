#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "eval/" --csv --log-file gpurun_out/r2c_c2_launches.csv python scripts/c2_eval.py 3 > gpurun_out/c2_eval.log 2>&1
tail -2 gpurun_out/c2_eval.log
python - <<'P'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2c_c2_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
tot = 0.0
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    us = v / 1e3 if u in ('ns', 'nsecond') else v
    tot += us
    print(f"{us:8.2f} us  {r[ki][:110]}")
print('launches', len(rows) - 1, 'sum us', tot)
P
