#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_sgpr.py tests/test_gpu_svgp.py -x -q 2>&1 | tail -3
python scripts/mm_probe.py 1024 2>&1 | tail -6
GGP_MM64_NO_FOLD=1 python scripts/mm_probe.py 1024 2>&1 | tail -5
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/q_{tag}.json').read().strip().splitlines()[-1])
    print(tag,'ms/step',round(d['ms_per_step'],2),'breakdown',{k:round(x,2) for k,x in d['breakdown_ms_per_step'].items()}, 'launches/step', d['gpu_launches']/d['steps'], 'sm_mhz',d['clocks']['sm_mhz'])
except Exception as e:
    print(tag,'FAILED',e); print(open(f'gpurun_out/q_{tag}.err').read()[-1500:])
PY
}
run fold GGP_DUMMY=1
run nofold GGP_MM64_NO_FOLD=1
