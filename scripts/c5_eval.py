"""One chain-batched logp/dlogp evaluation of BASELINE configs[4] (Bernoulli SGPMC, N=2e5, D=16, M=512, C chains): the launch sequence
of one leapfrog, for an ncu launch list / CUDA-event timing."""
import sys, time
import torch
sys.path.insert(0, '.')
import ggp_b200
import ggp_b200.synthetic as syn
from ggp_b200.functions import sgpmc_logp_dlogp
dev = torch.device('cuda:0')
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = syn.config5_classification()
X, y, Z = (torch.tensor(c[k], device=dev) for k in ("X", "y", "Z"))
N, D = X.shape; M = Z.shape[0]
eng = ggp_b200.Engine.get(dev)
g = torch.Generator(device=dev).manual_seed(173)
x0 = torch.cat([0.1 * torch.randn(C, M, dtype=torch.float64, device=dev, generator=g), torch.full((C, D + 2), 1.0, dtype=torch.float64, device=dev)], dim=1)
for i in range(n):
    torch.cuda.synchronize(); t0 = time.time()
    lp, gv, gr = sgpmc_logp_dlogp(x0[:, :M], x0[:, M:], X, y, Z, likelihood="bernoulli", engine=eng)
    torch.cuda.synchronize(); print("eval %d: %.1f ms" % (i, 1e3 * (time.time() - t0)))
print(float(lp.sum()))
