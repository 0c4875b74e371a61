import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ggp_b200
dev = torch.device("cuda:0"); eng = ggp_b200.Engine.get(dev)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
M, N = 1024, 16384 + 2560   # 1184 tiles = 8 x 148: no wave quantisation
C = torch.empty(M, N, dtype=torch.float64, device=dev)
for K in [128, 256, 512, 1024, 2048, 4096]:
    A = torch.randn(M, K, dtype=torch.float64, device=dev); B = torch.randn(N, K, dtype=torch.float64, device=dev)
    ms = timeit(lambda: eng.gemm_nt(A, B, C))
    per_tile_us = ms * 1e3 / 8
    print(f"K={K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.2f} TF/s  per tile {per_tile_us:.1f} us  per k-iter {per_tile_us/(K/32):.3f} us")
