#!/bin/bash
# developer loop on the GPU box: parity of both paths after the Q A backward (MN-major B operand on the sliced-integer path)
rm -f gpurun_out/headline_parity.json
timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_sgpr.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_headline_parity.py -q 2>&1 | tail -15
cat gpurun_out/headline_parity.json
bash scripts/quick_variants.sh -
