"""Developer tool: warm CUDA-event timings of the pieces of the m x m section at m = 1024 (blocked Cholesky + explicit inverse as
one graph replay, and the triangular-clipped 1024^3 products), for A/B runs with GGP_MM64_MAX_TILES / GGP_CHOL_LOOKAHEAD."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ggp_b200
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0"); eng = ggp_b200.Engine.get(dev)
g = torch.Generator(device=dev).manual_seed(0)
R = torch.randn(1, m, m + 64, dtype=torch.float64, device=dev, generator=g)
A = R @ R.transpose(1, 2) / m + torch.eye(m, dtype=torch.float64, device=dev)
L, Li, info = eng.chol(A)
print("info", info.tolist(), "chol err", float((L[0] @ L[0].T - A[0]).abs().max() / A[0].abs().max()),
      "inverse err", float((Li[0] @ L[0] - torch.eye(m, dtype=torch.float64, device=dev)).abs().max()))
def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
print("chol + inverse (incl. the python wrapper's copies): %.1f us" % timed(lambda: eng.chol(A)))
U = torch.triu(torch.randn(m, m, dtype=torch.float64, device=dev, generator=g))
F = torch.randn(m, m, dtype=torch.float64, device=dev, generator=g)
C = torch.zeros(m, m, dtype=torch.float64, device=dev)
for name, a, b, km, ref in (("full", F, F, 0, F @ F.T), ("A upper", U, F, 2, U @ F.T), ("A, B upper", U, U, 10, U @ U.T),
                            ("B upper", F, U, 8, F @ U.T)):
    eng.gemm_nt_ex(a, b, C, kmode=km)
    err = float((C - ref).abs().max() / ref.abs().max())
    print("product %-10s %.1f us  err %.1e" % (name, timed(lambda: eng.gemm_nt_ex(a, b, C, kmode=km)), err))
