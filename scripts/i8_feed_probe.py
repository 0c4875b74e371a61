"""Developer tool: time the sliced-integer GEMM launch alone (GGP_I8_EXP_TIME) under the experiment switches
GGP_I8_EXP_NOEPI / GGP_I8_EXP_SKIPA / GGP_I8_EXP_SKIPB, on a long-k shape where the drain is negligible."""
import os, sys
os.environ["GGP_I8_EXP_TIME"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ggp_b200
dev = torch.device("cuda:0"); eng = ggp_b200.Engine.get(dev)
mm, nn, kk = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1024, 9472, 4096)
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(mm, kk, dtype=torch.float64, device=dev, generator=g); B = torch.rand(nn, kk, dtype=torch.float64, device=dev, generator=g)
eng.gemm_nt_i8(A, B)
