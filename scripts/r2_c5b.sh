#!/bin/bash
python scripts/c5_eval.py 64 3 2>&1 | tail -3
python scripts/c5_eval.py 8 3 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_svgp.py tests/test_gpu_models.py tests/test_gpu_sgpr.py -x -q 2>&1 | tail -3
