"""Developer tool: clock64 timeline of one 64 x 64 diagonal-block factorisation (GGP_POTF2_TIMELINE=1)."""
import os, sys
os.environ["GGP_POTF2_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ggp_b200
dev = torch.device("cuda:0"); eng = ggp_b200.Engine.get(dev)
g = torch.Generator(device=dev).manual_seed(0)
R = torch.randn(1, 256, 300, dtype=torch.float64, device=dev, generator=g)
A = R @ R.transpose(1, 2) + 300 * torch.eye(256, dtype=torch.float64, device=dev)
eng.chol(A)      # reserves the handle for m = 256
L, Li, info = eng.chol(A)
print("info", info.tolist(), "err", float((L[0] @ L[0].T - A[0]).abs().max() / A[0].abs().max()))
