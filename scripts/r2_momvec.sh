#!/bin/bash
# A/B: moments of the backward epilogue on the vector FP64 pipe (-DGGP_I8_MOM_VEC) against the DMMA route
V=generalised-gaussian-processes_b200/libggp_b200_momvec.so
GGP_B200_LIB=$V timeout 600 python -m pytest tests/test_gpu_i8.py tests/test_gpu_headline_parity.py -x -q -k "not full_n" 2>&1 | tail -3
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/q_{tag}.json').read().strip().splitlines()[-1])
    print(tag,'ms/step',round(d['ms_per_step'],2),'breakdown',{k:round(x,2) for k,x in d['breakdown_ms_per_step'].items()}, 'sm_mhz',d['clocks']['sm_mhz'],'parity',{k:float('%.2e'%v) for k,v in d['parity_at_headline']['vs_long_double']['fp64_i8'].items()})
except Exception as e:
    print(tag,'FAILED',e); print(open(f'gpurun_out/q_{tag}.err').read()[-1500:])
PY
}
run dmma GGP_DUMMY=1
run vec GGP_B200_LIB=$V
run dmma2 GGP_DUMMY=1
run vec2 GGP_B200_LIB=$V
