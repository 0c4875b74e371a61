#!/bin/bash
# fragment-layout epilogue: serial (default) against overlapped with the next tile's MMAs (GGP_I8_SERIAL_EPI=0)
for e in GGP_I8_SERIAL_EPI=0; do
  echo "== [$e]"
  env $e GGP_I8_TIMELINE=2 python scripts/prof_one_eval_i8.py 131072 2>&1 | grep -A14 "epilogue 2" | grep "tile  [3-6]" | tail -4 | cut -c1-220
done
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
run serial GGP_DUMMY=1
run overlap GGP_I8_SERIAL_EPI=0
run serial2 GGP_DUMMY=1
run overlap2 GGP_I8_SERIAL_EPI=0
