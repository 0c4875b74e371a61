#!/bin/bash
# developer A/B on the GPU box: for each library variant ("-" = the default build) a short headline bench
for v in "$@"; do
  if [ "$v" = "-" ]; then unset GGP_B200_LIB; else export GGP_B200_LIB=$PWD/generalised-gaussian-processes_b200/libggp_b200_$v.so; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/q_bench_$v.json 2> gpurun_out/q_bench_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/q_bench_{v}.json').read().strip().splitlines()[-1])
    print('variant',v,'ms/step',round(d['ms_per_step'],2),'e2e ms',round(d['e2e']['ms_per_step'],2),'breakdown',{k:round(x,2) for k,x in d['breakdown_ms_per_step'].items()},'sm_mhz',d['clocks']['sm_mhz'],'Z err',float('%.2e'%d['parity_at_headline']['vs_long_double']['fp64_i8']['Z']))
except Exception as e:
    print('FAILED',v,e); print(open(f'gpurun_out/q_bench_{v}.err').read()[-800:])
PY
done
