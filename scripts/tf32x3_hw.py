"""Hardware check of the 3xTF32 split-precision contraction (design study for GGP_PREC_TF32X3, DESIGN.md 4b'): the same evaluation as
scripts/tf32x3_emulation.py, but the three dense contractions run on the B200's tensor cores in TF32 with FP32 accumulation
(torch.matmul on float32 operands with allow_tf32 = True: cuBLAS TF32 kernels -- the numerics class of tcgen05.mma kind::tf32), each
FP64 operand split as x ~= hi + lo (two TF32-representable fp32 words) and x y ~= hi hi' + hi lo' + lo hi'.  Everything else (kernel
tiles, m x m section, epilogues) stays FP64, exactly as a split-precision build of this library would do.  Reference: the float64
CPU oracle.  Developer study script (not product; torch.matmul is used as the measuring instrument here)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sgpr as osgpr            # noqa: E402
from oracle.kernels import ard_kernel       # noqa: E402
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True


def tf32_rn(x32):
    i = x32.view(torch.int32)
    r = i + 0xFFF + ((i >> 13) & 1)
    return (r & ~0x1FFF).view(torch.float32)


def split2(x64):
    hi = tf32_rn(x64.to(torch.float32))
    lo = tf32_rn((x64 - hi.double()).to(torch.float32))
    return hi, lo


def mm3(A64, B64):
    """A B^T with both operands split in two TF32 words, three tensor-core products, FP32 accumulation inside each, FP64 sum of the three."""
    ah, al = split2(A64)
    bh, bl = split2(B64)
    return (al @ bh.T).double() + (ah @ bl.T).double() + (ah @ bh.T).double()


def relerr(a, b):
    return float((a.cpu() - b).norm() / b.norm())


def run(N, M, D=8, jitter=1e-6, s2v=0.1, chunk=8192, seed=0, which=("trmm", "syrk", "bwd")):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(N, D, generator=g, dtype=torch.float64)
    w1, w2 = torch.randn(D, generator=g, dtype=torch.float64), torch.randn(D, generator=g, dtype=torch.float64)
    y = torch.sin(X @ w1) + 0.5 * (X @ w2) + 0.1 * torch.randn(N, generator=g, dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    Z = X[torch.randperm(N, generator=g)[:M]].clone()
    ell = torch.full((D,), math.sqrt(D), dtype=torch.float64)
    sf2, s2 = torch.tensor(1.0, dtype=torch.float64), torch.tensor(s2v, dtype=torch.float64)
    Fo, go, jit = osgpr.sgpr_bound_and_grads_chunked(X, y, Z, ell, sf2, s2, jitter_policy=jitter, normalize="none", chunk=chunk)
    Xd, yd, Zd, elld = X.to(dev), y.to(dev), Z.to(dev), ell.to(dev)
    sf2d, s2d = sf2.to(dev), s2.to(dev)
    I = torch.eye(M, dtype=torch.float64, device=dev)
    Kzz = ard_kernel(Zd, Zd, elld, sf2d)
    L = torch.linalg.cholesky(Kzz + jit * I)
    Linv = torch.linalg.solve_triangular(L, I, upper=False)
    S = torch.zeros(M, M, dtype=torch.float64, device=dev); b = torch.zeros(M, dtype=torch.float64, device=dev)
    yty = yd @ yd; trc = torch.zeros((), dtype=torch.float64, device=dev)
    Ats = []
    for i0 in range(0, N, chunk):
        Kc = ard_kernel(Xd[i0:i0 + chunk], Zd, elld, sf2d)
        At = mm3(Kc, Linv) if "trmm" in which else Kc @ Linv.T
        Ats.append(At)
        trc += (sf2d - (At * At).sum(1)).sum()
        b += At.T @ yd[i0:i0 + chunk]
        AtT = At.T.contiguous()
        S += mm3(AtT, AtT) if "syrk" in which else AtT @ AtT.T
    S = 0.5 * (S + S.T)
    Bm = I + S / s2d
    LB = torch.linalg.cholesky(Bm)
    c = torch.linalg.solve_triangular(LB, b[:, None], upper=False)[:, 0] / s2d
    F = (-0.5 * N * osgpr.LOG2PI - 0.5 * N * torch.log(s2d) - torch.log(torch.diagonal(LB)).sum() - 0.5 * (yty / s2d - c @ c) - 0.5 * trc / s2d)
    LBinv = torch.linalg.solve_triangular(LB, I, upper=False)
    Binv = LBinv.T @ LBinv
    beta = Binv @ b
    PA = (I - Binv) / s2d - torch.outer(beta, beta) / s2d ** 3
    Q = Linv.T @ PA                      # the well-conditioned backward form of this library: dF/dKzx = Q A + u y^T
    u = Linv.T @ beta / s2d ** 2
    Gbar = Bm + Binv - 2.0 * I + torch.outer(beta, beta) / s2d ** 2
    Gzz = -0.5 * Linv.T @ Gbar @ Linv
    r = torch.zeros(M, dtype=torch.float64, device=dev); Qm = torch.zeros(M, D, dtype=torch.float64, device=dev); T = torch.zeros(M, D, dtype=torch.float64, device=dev)
    for ci, i0 in enumerate(range(0, N, chunk)):
        Xc, yc = Xd[i0:i0 + chunk], yd[i0:i0 + chunk]
        Kc = ard_kernel(Zd, Xc, elld, sf2d)
        A = Ats[ci]                      # [n, M]
        G = mm3(Q, A) if "bwd" in which else Q @ A.T
        W = (G + torch.outer(u, yc)) * Kc
        r += W.sum(1); Qm += W @ Xc; T += W @ (Xc * Xc)
    V = Gzz * Kzz
    rv = V.sum(1); Qv = V @ Zd; Tv = V @ (Zd * Zd)
    d_ell = ((Zd * Zd * r[:, None] - 2 * Zd * Qm + T).sum(0) + (Zd * Zd * rv[:, None] - 2 * Zd * Qv + Tv).sum(0)) / elld ** 3
    d_sf2 = (r.sum() + rv.sum()) / sf2d - N / (2.0 * s2d)
    d_Z = ((Qm - Zd * r[:, None]) + 2.0 * (Qv - Zd * rv[:, None])) / elld ** 2
    cond = float(torch.linalg.cond((Kzz + jit * I).cpu()))
    print(f"N={N} M={M} jitter={jitter} cond(Kzz)={cond:.1e} TF32x3 on: {','.join(which) or 'nothing (FP64 check)'}")
    print(f"   rel.err  F {abs(float(F) - float(Fo)) / abs(float(Fo)):.2e}  d_ell {relerr(d_ell, go['ell']):.2e}  "
          f"d_sf2 {abs(float(d_sf2) - float(go['sf2'])) / abs(float(go['sf2'])):.2e}  d_Z {relerr(d_Z, go['Z']):.2e}", flush=True)


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    print(torch.cuda.get_device_name(0))
    run(N, M, which=())
    run(N, M, which=("syrk",))
    run(N, M, which=("trmm", "syrk"))
    run(N, M, which=("trmm", "syrk", "bwd"))
    run(N, M, jitter=1e-4, which=("trmm", "syrk", "bwd"))
    run(N, 1024, which=("trmm", "syrk", "bwd"))
