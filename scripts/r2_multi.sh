#!/bin/bash
# multi-GPU check on one box: sharded-vs-unsharded agreement, then the bench line at N ranks
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 \
  > gpurun_out/r2g_bench_${N}gpu.json 2> gpurun_out/r2g_bench_${N}gpu.err
tail -3 gpurun_out/r2g_bench_${N}gpu.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/r2g_bench_{n}gpu.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'ms/step',round(d['ms_per_step'],2),'evals/s',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'breakdown',{k:round(x,2) for k,x in d['breakdown_ms_per_step'].items()})
print('parity',{k:float('%.2e'%v) for k,v in d['parity_at_headline']['vs_long_double']['fp64_i8'].items()})
print('dmma',round(d['fp64_dmma_path']['ms_per_step'],2),'c5',json.dumps(d.get('config5_hmc'))[:600])
PY
