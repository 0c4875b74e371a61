import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, ggp_b200
from helpers import make_problem
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
X, y, Z, th = make_problem(rows, 1024, 8, seed=1)
dev = torch.device("cuda:0"); eng = ggp_b200.Engine.get(dev, precision="fp64_i8")
X, y, Z, th = X.to(dev), y.to(dev), Z.to(dev), th.to(dev)
for _ in range(2):
    out = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-6)
torch.cuda.synchronize()
print(out["bound"].item())
