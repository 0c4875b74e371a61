"""Generate tests/golden/c4_headline_ld_{without,with}_replacement.npz: the bound + gradient of BASELINE configs[3] at FULL size
(N = 1e6, D = 8, M = 1024, trained-like theta) evaluated in x87 long double by oracle/hp (about 12 minutes per case on 8 cores).
Only the outputs are stored (8203 doubles); the inputs are regenerated from the seeded generator (ggp_b200.synthetic.config4_large).
The jitter is the level the float64 ladder (oracle.linalg.psd_safe_cholesky, same rule as the CUDA Cholesky) settles on."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ggp_b200.synthetic as syn  # noqa: E402
from oracle import hp  # noqa: E402
from oracle.kernels import ard_kernel  # noqa: E402
from oracle.linalg import psd_safe_cholesky  # noqa: E402

D = 8
for wr in [a == "with" for a in (sys.argv[1:] or ["without", "with"])]:
    c = syn.config4_large(with_replacement=wr)
    th = syn.theta_trained_like(D)
    Zt, tht = torch.tensor(c["Z"]), torch.tensor(th)
    _, jit = psd_safe_cholesky(ard_kernel(Zt, Zt, tht[:D], tht[D]), "gpytorch")
    t0 = time.time()
    F, g = hp.bound_grad(c["X"], c["y"], c["Z"], th, jit, "ld")
    name = f"c4_headline_ld_{'with' if wr else 'without'}_replacement.npz"
    np.savez(os.path.join(ROOT, "tests", "golden", name), F=F, d_ell=g["ell"], d_sf2=g["sf2"], d_s2=g["s2"], d_Z=g["Z"], jitter=jit,
             theta=th, N=c["X"].shape[0], x_checksum=float(c["X"].sum()), z_idx_head=c["Z_idx"][:16])
    print(name, "F", F, "jitter", jit, "seconds", time.time() - t0, flush=True)
