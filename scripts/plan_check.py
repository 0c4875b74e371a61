"""Developer check: the execution plans of the sliced-integer path and the DMMA path against the oracle and each other."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, ggp_b200
from helpers import make_problem, relerr
from oracle import sgpr as osgpr
N, M, D, jit = (int(a) for a in sys.argv[1:4]) + (1e-4,) if len(sys.argv) > 3 else (70001, 256, 4, 1e-4)
X, y, Z, th = make_problem(N, M, D, seed=77)
dev = torch.device("cuda:0")
Fo, go, _ = osgpr.sgpr_bound_and_grads_chunked(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=jit, normalize="none")
Fa, ga = osgpr.sgpr_grads_closed_form(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=jit, normalize="none") if N <= 80000 else (None, None)
plans = {"i8 headline": dict(precision="fp64_i8"), "i8 no cache": dict(precision="fp64_i8", tile_cache_mib=0),
         "i8 chunk 4096": dict(precision="fp64_i8", chunk_rows=4096), "dmma": dict(precision="fp64")}
res = {}
for name, kw in plans.items():
    out = ggp_b200.Engine.get(dev, **kw).sgpr_eval(X, y, Z, th, jitter_policy=jit)
    g = out["grad"][0].cpu()
    res[name] = out
    print(f"{name:14s} vs oracle: bound {relerr(out['bound'], Fo):.2e} ell {relerr(g[:D], go['ell']):.2e} sf2 {relerr(g[D], go['sf2']):.2e} "
          f"s2 {relerr(g[D+1], go['s2']):.2e} Z {relerr(g[D+2:].view(M, D), go['Z']):.2e}")
if ga is not None:
    print("oracle chunked vs closed form: Z", relerr(go["Z"], ga["Z"]), "ell", relerr(go["ell"], ga["ell"]))
for name in plans:
    print(f"{name:14s} vs dmma: grad {relerr(res[name]['grad'], res['dmma']['grad']):.2e}   vs i8 headline: {relerr(res[name]['grad'], res['i8 headline']['grad']):.2e}")
