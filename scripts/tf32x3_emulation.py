"""CPU emulation of the 3xTF32 split-precision contraction numerics (design study for GGP_PREC_TF32X3).

Models tcgen05.mma kind::tf32: operands are fp32 words whose low 13 mantissa bits are ignored, products are exact,
the accumulator is fp32 (TMEM).  Two accumulator-rounding models: "rn" (round to nearest) and "rz" (truncate after every
k=8 instruction) -- the hardware behaviour has to be probed on the box; this script tells us what each would cost.

x ~= hi + lo, hi = tf32(x), lo = tf32(x - hi);  x*y ~= hi*hi' + hi*lo' + lo*hi'   (lo*lo' dropped, 2^-22 relative)

Usage: python scripts/tf32x3_emulation.py [N] [M] [ksub]
"""
import os
import sys
import math
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sgpr as osgpr            # noqa: E402  (developer study script, not product)
from oracle.kernels import ard_kernel       # noqa: E402

torch.set_num_threads(8)


def tf32_trunc(x32):
    return (x32.view(torch.int32) & ~0x1FFF).view(torch.float32)


def tf32_rn(x32):
    i = x32.view(torch.int32)
    r = i + 0xFFF + ((i >> 13) & 1)
    return (r & ~0x1FFF).view(torch.float32)


def split2(x64):
    hi = tf32_rn(x64.to(torch.float32))
    lo = tf32_trunc((x64 - hi.double()).to(torch.float32))
    return hi, lo


def split3(x64):
    hi = tf32_rn(x64.to(torch.float32))
    r = x64 - hi.double()
    mid = tf32_rn(r.to(torch.float32))
    lo = tf32_trunc((r - mid.double()).to(torch.float32))
    return hi, mid, lo


def rz32(x64):
    f = x64.to(torch.float32)
    over = f.double().abs() > x64.abs()
    return torch.where(over, torch.nextafter(f, torch.zeros_like(f)), f)


def mm_terms(terms, ksub, mode):
    """sum over (A,B) pairs in `terms` of A @ B^T with fp32 accumulation over blocks of ksub, fp64 across blocks."""
    K = terms[0][0].shape[1]
    out = torch.zeros(terms[0][0].shape[0], terms[0][1].shape[0], dtype=torch.float64)
    for k0 in range(0, K, ksub):
        if mode == "rn":
            acc = torch.zeros(out.shape, dtype=torch.float32)
            for A, B in terms:
                acc += A[:, k0:k0 + ksub] @ B[:, k0:k0 + ksub].T
            out += acc.double()
        elif mode == "rz":
            acc = torch.zeros(out.shape, dtype=torch.float32)
            for kk in range(k0, min(k0 + ksub, K), 8):
                for A, B in terms:
                    blk = A[:, kk:kk + 8].double() @ B[:, kk:kk + 8].double().T
                    acc = rz32(acc.double() + blk)
            out += acc.double()
        else:  # exact accumulation of the split products
            for A, B in terms:
                out += A[:, k0:k0 + ksub].double() @ B[:, k0:k0 + ksub].double().T
    return out


def mm3(A64, B64, ksub, mode, a_parts=2, b_parts=2):
    As = split3(A64) if a_parts == 3 else split2(A64)
    Bs = split3(B64) if b_parts == 3 else split2(B64)
    terms = []
    for i, a in enumerate(As):
        for j, b in enumerate(Bs):
            if i + j <= max(a_parts, b_parts) - 1:
                terms.append((a, b))
    terms = terms[::-1]   # small terms first
    return mm_terms(terms, ksub, mode)


def relerr(a, b):
    return float((a - b).norm() / b.norm())


def run(N=32768, M=256, D=8, ksub=256, mode="rn", trmm="tf32x3", linv_parts=2, jitter=1e-6, s2v=0.1, chunk=8192, seed=0,
        bwd="tf32x3"):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(N, D, generator=g, dtype=torch.float64)
    w1, w2 = torch.randn(D, generator=g, dtype=torch.float64), torch.randn(D, generator=g, dtype=torch.float64)
    y = torch.sin(X @ w1) + 0.5 * (X @ w2) + 0.1 * torch.randn(N, generator=g, dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    Z = X[torch.randperm(N, generator=g)[:M]].clone()
    ell = torch.full((D,), math.sqrt(D), dtype=torch.float64)
    sf2, s2 = torch.tensor(1.0, dtype=torch.float64), torch.tensor(s2v, dtype=torch.float64)
    Fo, go, jit = osgpr.sgpr_bound_and_grads_chunked(X, y, Z, ell, sf2, s2, jitter_policy=jitter, normalize="none", chunk=chunk)

    I = torch.eye(M, dtype=torch.float64)
    Kzz = ard_kernel(Z, Z, ell, sf2)
    L = torch.linalg.cholesky(Kzz + jit * I)
    Linv = torch.linalg.solve_triangular(L, I, upper=False)
    S = torch.zeros(M, M, dtype=torch.float64); b = torch.zeros(M, dtype=torch.float64)
    yty = y @ y; trc = torch.zeros((), dtype=torch.float64)
    for i0 in range(0, N, chunk):
        Kc = ard_kernel(X[i0:i0 + chunk], Z, ell, sf2)            # [n, M]
        if trmm == "fp64":
            At = Kc @ Linv.T
        else:
            At = mm3(Kc, Linv, M, mode, 2, linv_parts)            # [n, M] = Kc Linv^T, fp32 accumulate over k = M
        At32 = At.to(torch.float32).double() if trmm != "fp64" else At
        trc += (sf2 - (At32 * At32).sum(1)).sum()                 # epilogue: row norms in fp64
        b += At32.T @ y[i0:i0 + chunk]                            # epilogue: A y in fp64 from the fp32 tile
        S += mm3(At32.T.contiguous(), At32.T.contiguous(), ksub, mode)
    S = 0.5 * (S + S.T)
    Bm = I + S / s2
    LB = torch.linalg.cholesky(Bm)
    c = torch.linalg.solve_triangular(LB, b[:, None], upper=False)[:, 0] / s2
    F = (-0.5 * N * osgpr.LOG2PI - 0.5 * N * torch.log(s2) - torch.log(torch.diagonal(LB)).sum()
         - 0.5 * (yty / s2 - c @ c) - 0.5 * trc / s2)
    F_trS = (-0.5 * N * osgpr.LOG2PI - 0.5 * N * torch.log(s2) - torch.log(torch.diagonal(LB)).sum()
             - 0.5 * (yty / s2 - c @ c) - 0.5 * (N * sf2 - torch.trace(S)) / s2)
    LBinv = torch.linalg.solve_triangular(LB, I, upper=False)
    Binv = LBinv.T @ LBinv
    beta = Binv @ b
    PA = (I - Binv) / s2 - torch.outer(beta, beta) / s2 ** 3
    P = Linv.T @ PA @ Linv
    u = Linv.T @ beta / s2 ** 2
    Gbar = Bm + Binv - 2.0 * I + torch.outer(beta, beta) / s2 ** 2
    Gzz = -0.5 * Linv.T @ Gbar @ Linv
    r = torch.zeros(M, dtype=torch.float64); Q = torch.zeros(M, D, dtype=torch.float64); T = torch.zeros(M, D, dtype=torch.float64)
    for i0 in range(0, N, chunk):
        Xc, yc = X[i0:i0 + chunk], y[i0:i0 + chunk]
        Kc = ard_kernel(Z, Xc, ell, sf2)                          # [M, n]
        if bwd == "fp64":
            G = P @ Kc
        else:
            G = mm3(P, Kc.T.contiguous(), M, mode)
        W = (G + torch.outer(u, yc)) * Kc
        r += W.sum(1); Q += W @ Xc; T += W @ (Xc * Xc)
    V = Gzz * Kzz
    rv = V.sum(1); Qv = V @ Z; Tv = V @ (Z * Z)
    d_ell = ((Z * Z * r[:, None] - 2 * Z * Q + T).sum(0) + (Z * Z * rv[:, None] - 2 * Z * Qv + Tv).sum(0)) / ell ** 3
    d_sf2 = (r.sum() + rv.sum()) / sf2 - N / (2.0 * s2)
    d_Z = ((Q - Z * r[:, None]) + 2.0 * (Qv - Z * rv[:, None])) / ell ** 2
    d_s2 = (-N / (2 * s2) + 0.5 * (Binv * S).sum() / s2 ** 2 + 0.5 * yty / s2 ** 2
            - (b @ beta) / s2 ** 3 + (beta @ S @ beta) / (2 * s2 ** 4) + trc / (2 * s2 ** 2))
    print(f"N={N} M={M} ksub={ksub} mode={mode} trmm={trmm} linv_parts={linv_parts} bwd={bwd} jitter={jitter} s2={s2v} "
          f"cond(Kzz)~{float(torch.linalg.cond(Kzz + jit * I)):.1e}")
    print(f"   F={float(Fo):.6f}  rel.err F {abs(float(F - Fo)) / abs(float(Fo)):.2e} (trace from tr S: {abs(float(F_trS - Fo)) / abs(float(Fo)):.2e})"
          f"  d_ell {relerr(d_ell, go['ell']):.2e} d_sf2 {abs(float(d_sf2 - go['sf2'])) / abs(float(go['sf2'])):.2e} "
          f"d_s2 {abs(float(d_s2 - go['s2'])) / abs(float(go['s2'])):.2e} d_Z {relerr(d_Z, go['Z']):.2e}")


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    for kw in [dict(mode="rn", ksub=8192), dict(mode="rn", ksub=512), dict(mode="rz", ksub=512), dict(mode="rz", ksub=128),
               dict(mode="rn", ksub=512, trmm="fp64"), dict(mode="rn", ksub=512, linv_parts=3),
               dict(mode="rn", ksub=512, trmm="fp64", bwd="fp64"),
               dict(mode="rn", ksub=512, jitter=1e-8), dict(mode="rn", ksub=512, jitter=1e-8, linv_parts=3)]:
        run(N=N, M=M, **kw)
