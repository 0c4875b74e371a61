"""Host-side restatement of the digit recoding of the sliced-integer path (csrc/gemm_i8.cuh: i8_exp_for, i8_digit_bytes): the
identities the kernels rely on, checked with exact integer arithmetic (no GPU)."""
import math
import numpy as np

C = 0x0000808080808080          # sum_{i>=1} 128 * 256^(6-i)


def exp_for(mx):
    if not mx > 0.0:
        return 0
    e = math.frexp(mx)[1] - 1 + 2            # ilogb(mx) + 2
    if math.ldexp(mx, -e) >= 0.498:
        e += 1
    return e


def digits_of(T):
    """i8_digit_bytes: bytes of (T + C) ^ C; byte 6 - i is digit i as int8."""
    V = ((T + C) ^ C) & ((1 << 64) - 1)
    return [((V >> (8 * (6 - i))) & 0xFF) - (256 if (V >> (8 * (6 - i))) & 0x80 else 0) for i in range(7)]


def test_exponent_rule_keeps_the_top_digit_in_int8():
    rs = np.random.RandomState(0)
    for mx in np.concatenate([np.exp(rs.uniform(-30, 30, 2000)), [1.0, 2.0, 0.498, 0.4979999, 1.99, 3.9999, 1e-300, 1e300]]):
        e = exp_for(float(mx))
        v = math.ldexp(float(mx), -e)
        assert 0.124 < v < 0.498                 # never more than two bits of headroom
        for sign in (1, -1):
            T = int(round(sign * v * 2.0 ** 56))
            d = digits_of(T)
            assert all(-128 <= x <= 127 for x in d), (mx, e, d)


def test_digits_reconstruct_the_fixed_point_value_exactly():
    rs = np.random.RandomState(1)
    lim = int(0.498 * 2 ** 56)
    vals = [0, 1, -1, lim, -(2 ** 55) + 1, 127, 128, -128, -129, 2 ** 48 - 1, 2 ** 48, -(2 ** 48)]
    vals += [int(x) for x in rs.randint(-2 ** 55 + 1, lim, size=20000, dtype=np.int64)]
    for T in vals:
        d = digits_of(T)
        assert all(-128 <= x <= 127 for x in d)
        assert sum(x * 256 ** (6 - i) for i, x in enumerate(d)) == T


def test_level_sums_fit_int32_and_groups_fit_the_double_mantissa():
    """|d_i d_j| <= 2^14, l + 1 <= 7 pairs per level: exact int32 accumulation up to K = 16384 (I8_MAX_K); four (three) level sums
    combined with 8-bit shifts stay below 2^53 for K <= 4096 (any K <= 16384): one exact int -> double conversion per group."""
    for K, group in ((4096, 4), (16384, 3)):
        a = [(l + 1) * 2 ** 14 * K for l in range(7)]
        assert max(a) < 2 ** 31
        for l0 in range(0, 7, group):
            lv = a[l0:l0 + group]
            t = sum(v * 256 ** (len(lv) - 1 - k) for k, v in enumerate(lv))
            assert t < 2 ** 53, (K, l0, t)


def test_kept_pairs_cover_56_bits():
    pairs = [(i, j) for i in range(7) for j in range(7) if i + j <= 6]
    assert len(pairs) == 28
    # the most significant dropped level (i + j = 7) sits 2^-72 below the operand scales: 8 pairs of |d d| <= 2^14, K terms
    K = 1024
    dropped = 8 * 2 ** 14 * math.sqrt(K) * 2.0 ** (-8 * 9)
    full = K * (0.25 * 0.25) / 16            # typical sum |a||b| / 2^(ea+eb) for operands a few bits below their row scale
    assert dropped / full < 2.0 ** -50
