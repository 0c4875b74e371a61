#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_i8.py -x -q -k "cluster or host_rows or plans" 2>&1 | tail -3
GGP_CHOL_CLUSTER_INV=1 timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_r2d.py > gpurun_out/r2g_sanitize_memcheck.log 2>&1; tail -3 gpurun_out/r2g_sanitize_memcheck.log
