"""Host restatement of the sliced-integer GEMM of csrc/gemm_i8.cuh (no GPU): row exponents, 7 balanced radix-256 digits, exact
integer level sums over the pairs i + j <= 6, grouped recombination in float64.  Checks the accuracy class the GPU test asserts
(< 2e-15 of sum |a||b|) and the exactness assumptions of the kernel against a long-double reference."""
import math
import numpy as np

C = 0x0000808080808080


def exp_for(mx):
    if not mx > 0.0:
        return 0
    e = math.frexp(mx)[1] + 1                    # ilogb(mx) + 2
    return e + 1 if math.ldexp(mx, -e) >= 0.498 else e


def slice_rows(X):
    """-> digits [7][rows][k] int64, exponents [rows]  (k_slice_rows)."""
    ex = np.array([exp_for(float(np.abs(r).max())) for r in X])
    T = np.rint(np.ldexp(X, (56 - ex)[:, None])).astype(np.int64)       # rn(x 2^(56 - e)), |T| < 0.498 * 2^56
    V = (T + C) ^ C
    D = np.stack([((V >> (8 * (6 - i))) & 0xFF).astype(np.int64) for i in range(7)])
    D = np.where(D >= 128, D - 256, D)
    assert (np.sum(D * (256 ** (6 - np.arange(7)))[:, None, None], axis=0) == T).all()
    return D, ex


def sliced_gemm(A, B):
    Da, ea = slice_rows(A)
    Db, eb = slice_rows(B)
    K = A.shape[1]
    lv = [sum(Da[i] @ Db[l - i].T for i in range(l + 1)) for l in range(7)]       # exact level sums (int64 here, int32 on the GPU)
    assert max(int(np.abs(x).max()) for x in lv) < 2 ** 31
    if K <= 4096:    # two groups: levels 0..3 and 4..6
        t_hi = ((lv[0] * 256 + lv[1]) * 256 + lv[2]) * 256 + lv[3]
        t_lo = (lv[4] * 256 + lv[5]) * 256 + lv[6]
        assert int(np.abs(t_hi).max()) < 2 ** 53 and int(np.abs(t_lo).max()) < 2 ** 53
        acc = t_lo.astype(np.float64) * 2.0 ** -64
        acc = t_hi.astype(np.float64) * 2.0 ** -40 + acc
    else:            # three groups: 0..2, 3..5, 6
        g0 = (lv[0] * 256 + lv[1]) * 256 + lv[2]
        g1 = (lv[3] * 256 + lv[4]) * 256 + lv[5]
        assert int(np.abs(g0).max()) < 2 ** 53 and int(np.abs(g1).max()) < 2 ** 53
        acc = g1.astype(np.float64) * 2.0 ** -56 + lv[6].astype(np.float64) * 2.0 ** -64
        acc = g0.astype(np.float64) * 2.0 ** -32 + acc
    return np.ldexp(acc, ea[:, None] + eb[None, :])


def _operands(mm, nn, kk, seed):
    rs = np.random.RandomState(seed)
    A = rs.randn(mm, kk) * np.exp2(20 * rs.rand(mm, 1) - 10) * np.exp2(-8 * rs.rand(mm, kk))     # rows of very different magnitude
    B = np.exp(-6 * rs.rand(nn, kk))                                                              # kernel-like values in (0, 1]
    return A, B


def test_sliced_gemm_is_fp64_class():
    for (mm, nn, kk) in ((40, 24, 64), (33, 17, 1000), (16, 16, 4096), (8, 8, 16384)):
        A, B = _operands(mm, nn, kk, seed=mm + kk)
        Cs = sliced_gemm(A, B)
        ref = (A.astype(np.longdouble) @ B.astype(np.longdouble).T)
        scale = np.abs(A) @ np.abs(B).T
        err = float((np.abs(Cs.astype(np.longdouble) - ref) / scale).max())
        fp64 = float((np.abs((A @ B.T).astype(np.longdouble) - ref) / scale).max())
        assert err < 2e-15, (mm, nn, kk, err)
        assert err < 20 * max(fp64, 1e-17) + 4e-16          # same class as (here: better than or close to) a float64 matmul


def test_all_integer_digit_epilogue_matches_the_float_route():
    """The triangular multiply's epilogue builds the output digits from (t_hi 2^24 + t_lo) >> sh in integers; that must equal
    rn(value 2^(56 - eo)) computed the long way, for shifts of both signs."""
    rs = np.random.RandomState(3)
    for sh in (-5, -1, 0, 1, 7, 13, 24):
        t_hi = rs.randint(-2 ** 30, 2 ** 30, size=2000, dtype=np.int64) >> max(0, -sh)      # keep the result below 2^55
        t_lo = rs.randint(-2 ** 40, 2 ** 40, size=2000, dtype=np.int64)
        for th, tl in zip(t_hi.tolist(), t_lo.tolist()):
            U = th * 2 ** 24 + tl
            if sh > 0:
                lo = (tl + (1 << (sh - 1))) >> sh
                fx = (th << (24 - sh)) + lo
                exact = (U + (1 << (sh - 1))) >> sh                                        # round half up, as the kernel
            else:
                fx = (th << (24 - sh)) + (tl << (-sh))
                exact = U << (-sh)
            assert fx == exact
