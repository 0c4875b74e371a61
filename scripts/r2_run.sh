#!/bin/bash
echo "== GPU tests (models + headline parity)"
rm -f gpurun_out/headline_parity.json
timeout 2400 python -m pytest tests/test_gpu_models.py tests/test_gpu_headline_parity.py -q 2>&1 | grep -v "^  \|^$\|Warning" | tail -30
echo "== bench (default flags)"
time (timeout 1500 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err); tail -5 gpurun_out/r2a_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2a_bench.json').read().strip().splitlines()[-1])
for k,v in d.items():
    print(k, json.dumps(v)[:900])
PY
