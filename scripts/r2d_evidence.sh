#!/bin/bash
# round-2d profile evidence: (a) launch list of the bench command, (b) ncu --set full of the backward GEMM (fragment-layout epilogue) and of
# the small-tile m x m kernel, raw pages as CSV
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_bench_N1e6_i8.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/r2d_ncu_bench.log 2>&1
tail -1 gpurun_out/r2d_ncu_bench.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:k_gemm_i8 -s 3 -c 3 -f -o gpurun_out/r2d_i8 python scripts/prof_one_eval_i8.py 16384 > gpurun_out/ncu_i8.log 2>&1
tail -1 gpurun_out/ncu_i8.log
ncu --set full --clock-control none --import-source on -k regex:k_mm64 -s 28 -c 2 -f -o gpurun_out/r2d_mm64 python scripts/prof_one_eval_i8.py 16384 > gpurun_out/ncu_mm64.log 2>&1
tail -1 gpurun_out/ncu_mm64.log
ncu -i gpurun_out/r2d_i8.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_i8_raw.csv 2>/dev/null
ncu -i gpurun_out/r2d_mm64.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_mm64_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep gpurun_out/r2d_*.csv
