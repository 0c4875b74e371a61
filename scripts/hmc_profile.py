"""Developer tool: one eager evaluation batch at the co2-shaped config (for an ncu launch list of the small-problem path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ggp_b200
import ggp_b200.synthetic as syn
from ggp_b200.functions import sgpr_vfe_logp_dlogp
dev = torch.device("cuda:0")
c = syn.config2_co2_shaped()
X, y, Z = (torch.tensor(c[k], device=dev) for k in ("X", "y", "Z"))
eng = ggp_b200.Engine.get(dev)
x = torch.zeros(4, 3, dtype=torch.float64, device=dev)
for _ in range(3):
    lp, g = sgpr_vfe_logp_dlogp(x, X, y, Z, engine=eng, group=False)
torch.cuda.synchronize()
print(lp.tolist())
