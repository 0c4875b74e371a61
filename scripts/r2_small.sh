#!/bin/bash
# small-problem path: one-launch factor + inverse for Mp <= 128; Cholesky-chain timeline at m = 1024; configs[0..2] legs
GGP_CHOL_TIMELINE=2 python scripts/mm_probe.py 1024 2 2>&1 | grep -v "^$" | tail -40
timeout 900 python -m pytest tests/test_gpu_sgpr.py tests/test_gpu_svgp.py tests/test_gpu_models.py tests/test_gpu_composite.py -x -q 2>&1 | tail -3
leg() {
env "$@" timeout 900 python - <<'P' 2>&1 | tail -4
import json, sys, torch
sys.path.insert(0, '.')
import bench, ggp_b200
dev = torch.device('cuda:0')
out = bench.hmc_leg(dev, ggp_b200.Engine, True)
print({k: (round(v['samples_per_s'], 1), round(v.get('ms_per_batched_eval', v.get('ms_per_batched_leapfrog', 0)), 4)) for k, v in
       dict(nuts=out['nuts_pymc3_defaults'], hmc_graph=out['fixed_length_hmc_L10']['cuda_graph']).items()})
P
}
leg GGP_DUMMY=1
leg GGP_CHOL_SMALL=0
leg GGP_CHOL_SMALL=0 GGP_MM64_MAX_TILES=0 GGP_CHOL_LOOKAHEAD=0
