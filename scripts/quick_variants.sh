#!/bin/bash
# developer A/B: short headline bench under a list of environment settings ("NAME=VAL NAME2=VAL2" per argument; "-" = defaults)
for v in "$@"; do
  if [ "$v" = "-" ]; then envs=""; else envs="$v"; fi
  env $envs timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hmc ${BENCH_EXTRA} > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
  echo "== variant: $v"
  python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
    print('ms/step',round(d['ms_per_step'],2),'evals/s',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'breakdown',{k:round(v,2) for k,v in d['breakdown_ms_per_step'].items()})
    f=d.get('fp64_dmma_path',{})
    print('   frac',round(d['roofline']['frac'],3),'sm_mhz',d['clocks']['sm_mhz'],'vs dmma: bound',f.get('rel_diff_of_bound_vs_headline_path'),'grad',f.get('rel_diff_of_grad_vs_headline_path'))
except Exception as e:
    print('FAILED',e); print(open('gpurun_out/q_bench.err').read()[-1500:])
PY
done
