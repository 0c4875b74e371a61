#!/bin/bash
# final-build evidence: launch list of the bench command; ncu --set full of the cluster-resident factorisation
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_bench_N1e6_i8.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-hmc --no-legs --no-dmma-leg > gpurun_out/r2h_ncu_bench.log 2>&1
tail -1 gpurun_out/r2h_ncu_bench.log | cut -c1-120
ncu --set full --clock-control none --import-source on -k regex:k_chol_cluster -s 2 -c 1 -f -o gpurun_out/r2h_cluster python scripts/prof_one_eval_i8.py 131072 > gpurun_out/ncu_cluster.log 2>&1
tail -1 gpurun_out/ncu_cluster.log
ncu -i gpurun_out/r2h_cluster.ncu-rep --page raw --csv > gpurun_out/r2h_ncu_cluster_raw.csv 2>/dev/null
ls -la gpurun_out/r2h_*
