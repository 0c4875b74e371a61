#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_sgpr.py -x -q -k "gemm or cholesky" 2>&1 | tail -2
GGP_CHOL_CLUSTER_N=16 timeout 300 python -m pytest tests/test_gpu_i8.py -x -q -k "cluster" 2>&1 | tail -2
for e in GGP_CHOL_CLUSTER_N=8 GGP_CHOL_CLUSTER_N=16 "GGP_CHOL_CLUSTER_N=16 GGP_CHOL_CLUSTER_INV=1"; do
  env $e timeout 300 python bench.py --rows 125000 --steps 20 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/q_thr.json 2> gpurun_out/q_thr.err
  python - "$e" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/q_thr.json').read().strip().splitlines()[-1])
print(sys.argv[1],'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),{k:round(x,2) for k,x in d['breakdown_ms_per_step'].items()},'launches/step', d['gpu_launches']/d['steps'],'sm_mhz',d['clocks']['sm_mhz'])
PY
done
