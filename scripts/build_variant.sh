#!/bin/bash
# developer A/B: build libggp_b200_<name>.so with extra nvcc flags (e.g. "-DGGP_I8_BKB=32 -DGGP_I8_STAGES=4"); run it with GGP_B200_LIB=...
name=$1; shift
d=generalised-gaussian-processes_b200
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC "$@" \
  -o $d/libggp_b200_$name.so $d/csrc/ggp_api.cu 2>&1 | grep -E "error|warning: v" ; ls -la $d/libggp_b200_$name.so
