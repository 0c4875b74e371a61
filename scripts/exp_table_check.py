"""Host check of the table-based exponential of csrc/kernel_tiles.cuh (exp_neg): emulates its FP64 steps (the two FMAs of the range
reduction in long double) and reports the error in ulp against a long-double exp over [-700, 0]."""
import numpy as np
from decimal import Decimal, getcontext
getcontext().prec = 50
ld = np.longdouble
ln2 = Decimal(2).ln()
T = [(ln2 * Decimal(j) / Decimal(64)).exp() for j in range(64)]
T_hi = np.array([np.float64(float(t)) for t in T])
T_lo = np.array([np.float64(float(t - Decimal(float(h)))) for t, h in zip(T, T_hi)])
L = ln2 / Decimal(64)
L_hi = np.float64(float(L)); L_lo = np.float64(float(L - Decimal(float(L_hi)))); INV = np.float64(float(Decimal(64) / ln2))
rs = np.random.RandomState(0)
x = np.concatenate([-np.abs(rs.randn(2_000_000)) * 8, -rs.rand(200_000) * 1e-3, -rs.rand(200_000) * 700, [0.0]])
k = np.rint(x * INV)
r = np.float64(ld(x) - ld(k) * ld(L_hi)); r = np.float64(ld(r) - ld(k) * ld(L_lo))
p = np.float64(1 / 720)
for c in (1 / 120, 1 / 24, 1 / 6, 0.5):
    p = p * r + c
p = p * r * r + r
ki = k.astype(np.int64)
res = np.ldexp(T_hi[ki & 63] + (T_lo[ki & 63] + T_hi[ki & 63] * p), (ki >> 6).astype(np.int32))
ref = np.exp(ld(x))
err = np.abs(ld(res) - ref) / np.spacing(np.float64(ref))
print(f"L_hi={float(L_hi)!r} L_lo={float(L_lo)!r} INV={float(INV)!r}  max |r|={np.abs(r).max():.6g}")
print(f"max error {float(err.max()):.3f} ulp, mean {float(err.mean()):.3f} ulp; exp(0) = {res[-1]!r}")
