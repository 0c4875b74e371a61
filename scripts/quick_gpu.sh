#!/bin/bash
# developer loop on the GPU box: i8 parity tests, then a short headline bench (no CPU baseline, no HMC leg)
timeout 300 python -m pytest tests/test_gpu_i8.py tests/test_gpu_sgpr.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hmc > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print('ms/step',d['ms_per_step'],'evals/s',d['value'],'e2e',d['e2e']['value'])
print('breakdown',{k:round(v,2) for k,v in d['breakdown_ms_per_step'].items()})
print('frac',d['roofline']['frac'],'clocks',d['clocks'])
f=d.get('fp64_dmma_path',{})
print('vs dmma: bound',f.get('rel_diff_of_bound_vs_headline_path'),'grad',f.get('rel_diff_of_grad_vs_headline_path'))
PY
