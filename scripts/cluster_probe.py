"""Developer tool: the cluster-resident Kzz factorisation ALONE (no tile build beside it), with and without the inverse in the launch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, ggp_b200
from helpers import make_problem
dev = torch.device("cuda:0")
eng = ggp_b200.Engine.get(dev, precision="fp64_i8")
X, y, Z, th = make_problem(4096, 1024, 8, seed=1)
Z, th = Z.to(dev), th.to(dev).unsqueeze(0)
eng.reserve(400000, 1024, 8, 1)
def run():
    eng.factor(Z, th, 1e-6, True, after_first_enqueue=lambda: None)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print(os.environ.get("GGP_CHOL_CLUSTER_INV"), os.environ.get("GGP_CHOL_CLUSTER_NO_INV"), os.environ.get("GGP_CHOL_CLUSTER"), "factor call: %.3f ms" % (e0.elapsed_time(e1) / 20))
