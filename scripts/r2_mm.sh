#!/bin/bash
python scripts/potf2_timeline.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_sgpr.py tests/test_gpu_i8.py tests/test_gpu_models.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
    print('ms/step',round(d['ms_per_step'],2),'evals/s',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'breakdown',{k:round(x,2) for k,x in d['breakdown_ms_per_step'].items()}, 'launches/step', d['gpu_launches']/d['steps'])
    r=d['roofline']; print('   frac',round(r['frac'],3),'sm_mhz',d['clocks']['sm_mhz'],'parity',{k:float('%.2e'%v) for k,v in d['parity_at_headline']['vs_long_double']['fp64_i8'].items()})
except Exception as e:
    print('FAILED',e); print(open('gpurun_out/q_bench.err').read()[-1500:])
PY
