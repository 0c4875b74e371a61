"""One batched logp/dlogp evaluation of BASELINE configs[1] (co2-shaped, 4 chains): the launch sequence a NUTS leapfrog replays."""
import sys
import torch
sys.path.insert(0, '.')
import ggp_b200
import ggp_b200.synthetic as syn
from ggp_b200.functions import sgpr_vfe_logp_dlogp
dev = torch.device('cuda:0')
c = syn.config2_co2_shaped()
X, y, Z = (torch.tensor(c[k], device=dev) for k in ("X", "y", "Z"))
eng = ggp_b200.Engine.get(dev)
x0 = torch.zeros(4, 3, dtype=torch.float64, device=dev)
x0[:, 0] = 0.69
x0 = x0 + 0.1 * torch.arange(4, dtype=torch.float64, device=dev).unsqueeze(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(n):
    if i == n - 1:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("eval")
    lp, g = sgpr_vfe_logp_dlogp(x0, X, y, Z, engine=eng, group=False)
    if i == n - 1:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
print(lp.tolist())
