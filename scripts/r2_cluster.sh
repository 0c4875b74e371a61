#!/bin/bash
# cluster-resident Kzz factorisation (+ inverse) next to the tile build (default) against: inverse as launches, the launch chain
timeout 300 python -m pytest tests/test_gpu_i8.py tests/test_gpu_headline_parity.py -x -q -k "not full_n" 2>&1 | tail -3
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg ${ROWS:+--rows $ROWS} > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/q_{tag}.json').read().strip().splitlines()[-1])
    par=d.get('parity_at_headline') or {}
    print(tag,'ms/step',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'breakdown',{k:round(x,2) for k,x in d['breakdown_ms_per_step'].items()}, 'launches/step', d['gpu_launches']/d['steps'], 'sm_mhz',d['clocks']['sm_mhz'],
          'parity',{k:float('%.2e'%v) for k,v in par['vs_long_double']['fp64_i8'].items()} if 'vs_long_double' in par else None)
except Exception as e:
    print(tag,'FAILED',e); print(open(f'gpurun_out/q_{tag}.err').read()[-1500:])
PY
}
run cluster_inv GGP_DUMMY=1
run cluster GGP_CHOL_CLUSTER_NO_INV=1
run chain GGP_CHOL_CLUSTER=0
ROWS=125000
run cluster_inv_shard GGP_DUMMY=1
run cluster_shard GGP_CHOL_CLUSTER_NO_INV=1
run chain_shard GGP_CHOL_CLUSTER=0
