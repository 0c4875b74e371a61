"""Developer tool (run under compute-sanitizer): the kernels that are new in round 2 on small shapes -- micro-blocked Cholesky,
predictive pass 1 + tiled covariance, SVGP gradients for the RQ tile, chain-batched SGPMC, the MN-major backward operand."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, ggp_b200
from helpers import make_problem
import ggp_b200.functions as F
dev = torch.device("cuda:0")
eng = ggp_b200.Engine.get(dev)
g = torch.Generator().manual_seed(0)
R = torch.randn(2, 200, 260, dtype=torch.float64, generator=g)
A = (R @ R.transpose(1, 2) + 260 * torch.eye(200, dtype=torch.float64)).to(dev)
L, Li, info = eng.chol(A)
print("chol", info.tolist(), float((L[1] @ L[1].T - A[1]).abs().max()))
X, y, Z, th = make_problem(900, 70, 3, seed=1)
eng.sgpr_predict_state(X, y, Z, th, jitter_policy=1e-6)
m, v, c = eng.sgpr_predict(torch.randn(2100, 3, dtype=torch.float64), Z, th, full_cov=True)
print("predict", float(m.abs().max()), tuple(c.shape))
erq = ggp_b200.Engine.get(dev, kernel="rq", kernel_param=1.3)
o = erq.svgp_eval(X[:500], y[:500], Z, torch.zeros(70, dtype=torch.float64), torch.eye(70, dtype=torch.float64), th, num_data=900)
print("svgp rq", float(o["value"][0]))
vv = 0.1 * torch.randn(3, 70, dtype=torch.float64, generator=g)
raw = torch.ones(3, 5, dtype=torch.float64)
lp, gv, gr = F.sgpmc_logp_dlogp(vv, raw, X.to(dev), (y > 0).double().to(dev), Z.to(dev), likelihood="bernoulli", engine=eng)
print("sgpmc", lp.tolist())
X2, y2, Z2, th2 = make_problem(3001, 130, 5, seed=2)
e8 = ggp_b200.Engine.get(dev, precision="fp64_i8", chunk_rows=1024)
e8.prefetch_min_rows = 1024
o8 = e8.sgpr_eval(X2, y2, Z2, th2, jitter_policy=1e-4)
print("i8", float(o8["bound"][0]), o8["path"])
torch.cuda.synchronize()
