# developer A/B: TMA-fed mainloop (default) vs LDGSTS mainloop (GGP_GEMM_TMA=0)
for v in 1 0; do
  echo "== GGP_GEMM_TMA=$v"
  export GGP_GEMM_TMA=$v
  timeout 120 python scripts/ksweep.py 2>&1 | tail -6
  timeout 300 python bench.py --rows 262144 --steps 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['breakdown_ms_per_step'])"
done
