#!/bin/bash
for e in "" "GGP_I8_F64_BN32=1" "GGP_I8_BN64=1"; do
  echo "== env [$e]: i8 parity tests"
  env $e timeout 600 python -m pytest tests/test_gpu_i8.py -x -q 2>&1 | tail -3
done
for e in "" "GGP_I8_BN64=1" ""; do
  echo "== env [$e]: bench"
  env $e timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
  python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
    print('ms/step',round(d['ms_per_step'],2),'evals/s',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'breakdown',{k:round(x,2) for k,x in d['breakdown_ms_per_step'].items()})
    r=d['roofline']; print('   frac',round(r['frac'],3),'of mix',round(r['frac_of_the_production_mma_mix'],3),'peak',round(r['peak']),'sm_mhz',d['clocks']['sm_mhz'],'parity',{k:float('%.2e'%v) for k,v in d['parity_at_headline']['vs_long_double']['fp64_i8'].items()})
except Exception as e:
    print('FAILED',e); print(open('gpurun_out/q_bench.err').read()[-1500:])
PY
done
