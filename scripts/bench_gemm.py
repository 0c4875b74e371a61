"""GEMM micro-benchmarks on the GPU box (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ggp_b200
dev = torch.device("cuda:0")
eng = ggp_b200.Engine.get(dev)

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for (mm, nn, kk) in [(8192, 8192, 8192), (1024, 16384, 1024), (1024, 1024, 16384), (4096, 4096, 1024)]:
    A = torch.randn(mm, kk, dtype=torch.float64, device=dev); B = torch.randn(nn, kk, dtype=torch.float64, device=dev)
    C = torch.empty(mm, nn, dtype=torch.float64, device=dev)
    ms = timeit(lambda: eng.gemm_nt(A, B, C))
    msc = timeit(lambda: torch.matmul(A, B.T, out=C))
    print(f"gemm {mm}x{nn}x{kk}: own {2*mm*nn*kk/ms/1e9:.2f} TF/s ({ms:.3f} ms)   cublas {2*mm*nn*kk/msc/1e9:.2f} TF/s")
for m in [128, 512, 1024]:
    R = torch.randn(1, m, m, dtype=torch.float64, device=dev)
    S = R @ R.transpose(1, 2) + m * torch.eye(m, dtype=torch.float64, device=dev)
    ms = timeit(lambda: eng.chol(S))
    msc = timeit(lambda: torch.linalg.cholesky(S))
    print(f"chol+inverse m={m}: own {ms:.3f} ms (incl. copies)  torch potrf only {msc:.3f} ms")
# structured variants (triangular A, symmetric output with split-K)
M, N = 1024, 16384
Linv = torch.tril(torch.randn(M, M, dtype=torch.float64, device=dev)); Kc = torch.randn(N, M, dtype=torch.float64, device=dev)
At = torch.empty(M, N, dtype=torch.float64, device=dev)
for km in [0, 1]:
    ms = timeit(lambda: eng.gemm_nt_ex(Linv, Kc, At, kmode=km))
    print(f"trmm-shape kmode={km}: {ms:.3f} ms  dense-equiv {2*M*N*M/ms/1e9:.2f} TF/s  algorithmic(tri) {M*N*M/ms/1e9:.2f} TF/s")
print("trmm err", float((At - Linv @ Kc.T).abs().max()))
Sp = torch.zeros(4, M, M, dtype=torch.float64, device=dev)
for (sym, sp) in [(0, 1), (1, 1), (1, 4), (0, 4)]:
    ms = timeit(lambda: eng.gemm_nt_ex(At, At, Sp, beta=1.0, sym=sym, splits=sp, split_stride=M * M))
    print(f"syrk-shape sym={sym} splits={sp}: {ms:.3f} ms  algorithmic(sym) {M*M*N/ms/1e9:.2f} TF/s")
