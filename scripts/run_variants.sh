# developer A/B: run the K sweep and a short bench against alternative builds of the library (GGP_B200_LIB override)
for v in "$@"; do
  echo "== $v"
  export GGP_B200_LIB=$PWD/generalised-gaussian-processes_b200/libggp_b200_$v.so
  python scripts/ksweep.py 2>&1 | tail -6
  python bench.py --rows 262144 --steps 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['breakdown_ms_per_step'])"
done
unset GGP_B200_LIB
