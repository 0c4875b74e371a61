"""Developer smoke script for the GPU box: building blocks + SGPR parity + probes (prints, never asserts)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ggp_b200
from oracle import sgpr as osgpr
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_problem, relerr

torch.manual_seed(0)
dev = torch.device("cuda:0")
eng = ggp_b200.Engine.get(dev)
print("device", torch.cuda.get_device_name(0))

# gemm
for (mm, nn, kk) in [(128, 128, 64), (200, 70, 33), (512, 384, 1000)]:
    A = torch.randn(mm, kk + (kk & 1), dtype=torch.float64, device=dev)[:, :kk]
    B = torch.randn(nn, kk + (kk & 1), dtype=torch.float64, device=dev)[:, :kk]
    C = eng.gemm_nt(A, B)
    print("gemm", mm, nn, kk, relerr(C, A @ B.T))

# chol
for m in [20, 64, 100, 500]:
    R = torch.randn(2, m, m, dtype=torch.float64, device=dev)
    S = R @ R.transpose(1, 2) + m * torch.eye(m, dtype=torch.float64, device=dev)
    L, Linv, info = eng.chol(S)
    Lt = torch.linalg.cholesky(S)
    print("chol", m, info.tolist(), relerr(torch.tril(L), Lt), relerr(Linv, torch.linalg.inv(Lt)))

# sgpr parity
for (N, M, D) in [(300, 20, 1), (1000, 100, 3), (3000, 260, 4)]:
    X, y, Z, th = make_problem(N, M, D, seed=N)
    out = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-6)
    Fo, go = osgpr.sgpr_grads_closed_form(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=1e-6, normalize="none")
    g = out["grad"][0].cpu()
    print("sgpr", N, M, D, "F", out["bound"].item(), Fo.item(), relerr(out["bound"], Fo),
          "ell", relerr(g[:D], go["ell"]), "sf2", relerr(g[D], go["sf2"]), "s2", relerr(g[D + 1], go["s2"]),
          "Z", relerr(g[D + 2:].view(M, D), go["Z"]), "info", out["info"].tolist(), out["info_b"].tolist())

print("dmma probe", eng.probe_dmma_peak(20000))
a = torch.randn(8192, 8192, dtype=torch.float64, device=dev); b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
for _ in range(2): (a @ b)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
print("cublas dgemm 8192^3 TF/s", 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
# own gemm rate
e0.record(); c2 = eng.gemm_nt(a, b); e1.record(); torch.cuda.synchronize()
e0.record(); c2 = eng.gemm_nt(a, b); e1.record(); torch.cuda.synchronize()
print("own dmma gemm 8192^3 TF/s", 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12, relerr(c2, a @ b.T))

# big eval timing
N, M, D = 131072, 1024, 8
X, y, Z, th = make_problem(N, M, D, seed=1)
X, y, Z, th = X.to(dev), y.to(dev), Z.to(dev), th.to(dev)
out = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-6)
torch.cuda.synchronize()
t0 = time.time(); out = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-6); torch.cuda.synchronize(); dt = time.time() - t0
print("eval N=131072 M=1024 D=8: %.1f ms, %.2f TF/s (4NM^2)" % (dt * 1e3, 4 * N * M * M / dt / 1e12), out["bound"].item())
