#!/bin/bash
rm -f gpurun_out/headline_parity.json
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$\|Warning" | tail -25
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
time (timeout 1500 python bench.py > gpurun_out/r2h_bench_1gpu.json 2> gpurun_out/r2h_bench_1gpu.err); tail -3 gpurun_out/r2h_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench_1gpu.json').read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],2),'evals/s',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'breakdown',{k:round(x,2) for k,x in d['breakdown_ms_per_step'].items()}, 'launches', d['gpu_launches'])
print('roofline frac',d['roofline']['frac'],'peak',d['roofline']['peak'],'clocks',d['clocks'])
print('parity',d['parity_at_headline'])
for k in ('fp64_dmma_path','cpu_baseline','config5_hmc','variants','svgp','config1','hmc'):
    print(k, json.dumps(d.get(k))[:800])
PY
