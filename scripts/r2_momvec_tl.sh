#!/bin/bash
V=generalised-gaussian-processes_b200/libggp_b200_momvec.so
W=generalised-gaussian-processes_b200/libggp_b200_diag.so
for lib in "$W" "$V"; do
  echo "== lib [$lib]"
  GGP_B200_LIB=$lib GGP_I8_TIMELINE=2 python scripts/prof_one_eval_i8.py 131072 2>&1 | grep -A14 "epilogue 2" | tail -15 | cut -c1-220
done
