#!/bin/bash
# native NUTS tree: tests + the configs[1] HMC/NUTS leg alone
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_sgpr.py -m gpu -q -k "nuts or pymc3" 2>&1 | grep -v "^  \|^$\|Warning" | tail -30
timeout 900 python - <<'P' 2>&1 | tail -5
import json, sys, torch
sys.path.insert(0, '.')
import bench, ggp_b200
dev = torch.device('cuda:0')
out = bench.hmc_leg(dev, ggp_b200.Engine, True)
json.dump(out, open('gpurun_out/r2c_hmc_leg.json', 'w'), indent=1)
print(json.dumps(out)[:3000])
P
