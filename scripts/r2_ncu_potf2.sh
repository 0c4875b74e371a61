#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:k_potf2_trti2 -s 40 -c 1 -o gpurun_out/r2_potf2 -f python scripts/prof_one_eval_i8.py 16384 > gpurun_out/ncu_potf2.log 2>&1
tail -3 gpurun_out/ncu_potf2.log
