"""Developer check under compute-sanitizer: one evaluation through each execution plan of the sliced-integer path and the DMMA path
at small shapes (ragged N, padded M)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, ggp_b200
from helpers import make_problem
dev = torch.device("cuda:0")
for (N, M, D) in ((66001, 200, 3), (3001, 130, 5)):
    X, y, Z, th = make_problem(N, M, D, seed=N)
    for kw in (dict(precision="fp64_i8"), dict(precision="fp64_i8", tile_cache_mib=0), dict(precision="fp64_i8", chunk_rows=4096),
               dict(precision="fp64")):
        eng = ggp_b200.Engine.get(dev, **kw)
        eng.prefetch_min_rows = 1024
        out = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-4)
        torch.cuda.synchronize()
        print(N, M, D, kw, float(out["bound"][0]), float(out["grad"].abs().max()))
