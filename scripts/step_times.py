"""Developer tool: per-step CUDA-event times of the resident bound+gradient evaluation at a given row count (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ggp_b200
import ggp_b200.synthetic as syn
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
dev = torch.device("cuda:0")
c = syn.config4_large(N=rows, D=8, M=1024)
X, y, Z = (torch.tensor(c[k]).to(dev) for k in ("X", "y", "Z"))
th = torch.tensor(syn.theta_trained_like(8)).to(dev)
eng = ggp_b200.Engine.get(dev, precision="fp64_i8")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(14)]
for i in range(3):
    eng.sgpr_eval(X, y, Z, th, jitter_policy="gpytorch", need_grad=True)
torch.cuda.synchronize()
import time
if "--profile" in sys.argv:
    eng.profile_read(); eng.profile_enable(True)
host = []
for i in range(13):
    ev[i].record()
    t0 = time.time()
    if "--noflush" not in sys.argv: flush.zero_()
    eng.sgpr_eval(X, y, Z, th, jitter_policy="gpytorch", need_grad=True)
    host.append(1e3 * (time.time() - t0))
ev[13].record(); torch.cuda.synchronize()
print({k: os.environ.get(k) for k in ("GGP_CHOL_CLUSTER_INV", "GGP_CHOL_CLUSTER_NO_INV")}, "gpu ms per step", [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(13)])
print("   host ms per step", [round(h, 2) for h in host])
