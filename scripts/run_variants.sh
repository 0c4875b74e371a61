for v in bk16s4 bk32s3 bk32s2; do
  echo "== $v"
  export GGP_B200_LIB=$PWD/generalised-gaussian-processes_b200/libggp_b200_$v.so
  python scripts/bench_gemm.py 2>&1 | tail -7
  python bench.py --rows 262144 --steps 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['breakdown_ms_per_step'])"
done
unset GGP_B200_LIB
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
