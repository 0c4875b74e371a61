"""Developer tool (run under compute-sanitizer): the kernels that are new in round 2d on small shapes -- the 64 x 64-tile products and the
look-ahead Cholesky step (Mp = 256), the one-launch factor + inverse (Mp = 64, 128, batched), the fragment-layout backward epilogue of the
sliced-integer path (one-launch plans), the lane-parallel Gauss-Hermite rows and the dm reduction of the SVGP / SGPMC path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, ggp_b200
from helpers import make_problem
import ggp_b200.functions as F
dev = torch.device("cuda:0")
eng = ggp_b200.Engine.get(dev)
g = torch.Generator().manual_seed(0)
for bsz, m in ((2, 200), (3, 100), (2, 40)):
    R = torch.randn(bsz, m, m + 60, dtype=torch.float64, generator=g)
    A = (R @ R.transpose(1, 2) + (m + 60) * torch.eye(m, dtype=torch.float64)).to(dev)
    L, Li, info = eng.chol(A)
    print("chol m=%d" % m, info.tolist(), float((L[-1] @ L[-1].T - A[-1]).abs().max()), float((Li[-1] @ L[-1] - torch.eye(m, dtype=torch.float64, device=dev)).abs().max()))
X, y, Z, th = make_problem(900, 70, 3, seed=1)
o = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-6)
print("sgpr m=70", float(o["bound"][0]))
X, y, Z, th = make_problem(2000, 200, 3, seed=3)
o = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-6)
print("sgpr m=200", float(o["bound"][0]))
vv = 0.1 * torch.randn(3, 70, dtype=torch.float64, generator=g)
raw = torch.ones(3, 5, dtype=torch.float64)
X, y, Z, th = make_problem(900, 70, 3, seed=1)
lp, gv, gr = F.sgpmc_logp_dlogp(vv, raw, X.to(dev), (y > 0).double().to(dev), Z.to(dev), likelihood="bernoulli", engine=eng)
print("sgpmc", lp.tolist())
X2, y2, Z2, th2 = make_problem(3001, 130, 5, seed=2)
e8 = ggp_b200.Engine.get(dev, precision="fp64_i8", chunk_rows=1024)
e8.prefetch_min_rows = 1024
o8 = e8.sgpr_eval(X2, y2, Z2, th2, jitter_policy=1e-4)
print("i8", float(o8["bound"][0]), o8["path"])
torch.cuda.synchronize()
