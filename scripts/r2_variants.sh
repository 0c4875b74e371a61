#!/bin/bash
# developer A/B on the GPU box: for each library variant ("-" = the default build) the i8 parity tests and a short headline bench
for v in "$@"; do
  if [ "$v" = "-" ]; then unset GGP_B200_LIB; else export GGP_B200_LIB=$PWD/generalised-gaussian-processes_b200/libggp_b200_$v.so; fi
  echo "== variant $v"
  timeout 600 python -m pytest tests/test_gpu_i8.py -x -q 2>&1 | tail -2
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/q_bench_$v.json 2> gpurun_out/q_bench_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/q_bench_{v}.json').read().strip().splitlines()[-1])
    print('ms/step',round(d['ms_per_step'],2),'evals/s',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'breakdown',{k:round(x,2) for k,x in d['breakdown_ms_per_step'].items()})
    r=d['roofline']; print('   frac',round(r['frac'],3),'of mix',round(r['frac_of_the_production_mma_mix'],3),'peak',round(r['peak']),'sm_mhz',d['clocks']['sm_mhz'],'parity',d['parity_at_headline'].get('vs_long_double'))
except Exception as e:
    print('FAILED',e); print(open(f'gpurun_out/q_bench_{v}.err').read()[-1500:])
PY
done
