#!/bin/bash
for v in "$@"; do
  if [ "$v" = "-" ]; then unset GGP_B200_LIB; else export GGP_B200_LIB=$PWD/generalised-gaussian-processes_b200/libggp_b200_$v.so; fi
  for e in "" "GGP_I8_EXP_NOEPI=1" "GGP_I8_EXP_NOEPI=1 GGP_I8_EXP_SKIPB=1" "GGP_I8_EXP_NOEPI=1 GGP_I8_EXP_SKIPA=1" "GGP_I8_EXP_NOEPI=1 GGP_I8_EXP_SKIPA=1 GGP_I8_EXP_SKIPB=1"; do
    echo "== variant $v env [$e]"
    env $e python scripts/i8_feed_probe.py 1024 9472 4096 2>&1 | grep "== i8 gemm"
    env $e python scripts/i8_feed_probe.py 1024 9472 1024 2>&1 | grep "== i8 gemm"
  done
done
python scripts/r2_probe.py
