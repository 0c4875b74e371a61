"""Developer tool: clock64 timeline of CTA 0 of the sliced-integer GEMM (GGP_I8_TIMELINE=1) on a backward-GEMM-like shape."""
import os, sys
os.environ["GGP_I8_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ggp_b200
dev = torch.device("cuda:0"); eng = ggp_b200.Engine.get(dev)
mm, nn, kk = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1024, 16384, 1024)
A = torch.randn(mm, kk, dtype=torch.float64, device=dev); B = torch.rand(nn, kk, dtype=torch.float64, device=dev)
for _ in range(2):
    C = eng.gemm_nt_i8(A, B)
print(float((C - A @ B.T).abs().max() / (A.abs() @ B.abs().T).max()))
