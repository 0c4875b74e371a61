#!/bin/bash
# profile evidence for profiles/: (a) launch list of the bench command, (b) ncu --set full of the three k_gemm_i8 roles and the tile build
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_N1e6.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/r2_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_ncu_bench.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:k_gemm_i8 -s 3 -c 3 -f -o gpurun_out/r2_i8 python scripts/prof_one_eval_i8.py 16384 > gpurun_out/ncu_i8.log 2>&1
tail -2 gpurun_out/ncu_i8.log
ncu --set full --clock-control none --import-source on -k regex:k_build_kc_i8 -s 1 -c 1 -f -o gpurun_out/r2_build python scripts/prof_one_eval_i8.py 16384 > gpurun_out/ncu_build.log 2>&1
tail -2 gpurun_out/ncu_build.log
compute-sanitizer --tool memcheck python scripts/sanitize_eval.py > gpurun_out/r2_sanitize.log 2>&1; tail -3 gpurun_out/r2_sanitize.log
