#!/bin/bash
# diagnostic: bench --rows 400000 with the inverse inside the cluster launch -- with and without the nvidia-smi sampler
for e in "GGP_CHOL_CLUSTER_INV=1" "GGP_CHOL_CLUSTER_INV=1 GGP_BENCH_NO_SAMPLER=1" "GGP_CHOL_CLUSTER_NO_INV=1 GGP_BENCH_NO_SAMPLER=1"; do
  env $e timeout 300 python bench.py --rows 400000 --steps 20 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/q_thr.json 2> gpurun_out/q_thr.err
  python - "$e" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/q_thr.json').read().strip().splitlines()[-1])
    print(sys.argv[1],'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'launches/step', d['gpu_launches']/d['steps'],'clocks',d['clocks'])
except Exception as ex:
    print(sys.argv[1],'FAILED',ex, open('gpurun_out/q_thr.err').read()[-800:])
PY
done
