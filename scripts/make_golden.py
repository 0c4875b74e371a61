"""Generate the committed golden vectors under tests/golden/ (run once, on CPU, in the build container).

The reference ships no fixtures and its evaluators (gpytorch / pymc3 / gpflow) are not installed here, so these
vectors come from (a) the DENSE textbook definitions in float64 (independent of any library's Woodbury / whitening
algebra) and (b) the real RNG objects: numpy's legacy global MT19937 stream and torch's DataLoader.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_problem  # noqa: E402
from oracle import sgpr, svgp, priors, sgpmc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_default_dtype(torch.float64)


def sgpr_case(name, N, M, D, seed, jitter):
    X, y, Z, th = make_problem(N, M, D, seed=seed)
    ell, sf2, s2 = th[:D], th[D], th[D + 1]
    F_dense = sgpr.sgpr_bound_dense(X, y, Z, ell, sf2, s2, jitter=jitter)
    # gradients of the DENSE definition by autograd (independent of the A/B algebra)
    p = [t.clone().requires_grad_(True) for t in (ell, sf2, s2, Z)]
    Fd = sgpr.sgpr_bound_dense(X, y, p[3], p[0], p[1], p[2], jitter=jitter)
    g = torch.autograd.grad(Fd, p)
    Xs = torch.tensor(np.random.RandomState(seed + 1).randn(37, D))
    # eval-mode predictive: gpytorch's (diagonal correction on the training rows too) and the plain-noise variant
    mean, cov = sgpr.sgpr_predict_dense(Xs, X, y, Z, ell, sf2, s2, jitter=jitter)
    mean_p, cov_p = sgpr.sgpr_predict_dense(Xs, X, y, Z, ell, sf2, s2, jitter=jitter, train_diag_correction=False)
    np.savez(os.path.join(OUT, name + ".npz"), X=X.numpy(), y=y.numpy(), Z=Z.numpy(), theta=th.numpy(), jitter=jitter,
             F_dense=F_dense.item(), g_ell=g[0].numpy(), g_sf2=g[1].item(), g_s2=g[2].item(), g_Z=g[3].numpy(),
             Xs=Xs.numpy(), pred_mean=mean.numpy(), pred_cov=cov.numpy(), pred_mean_plain=mean_p.numpy(), pred_cov_plain=cov_p.numpy())
    print(name, F_dense.item())


def sgpr_hp_case(name, N, M, D, seed, jitter):
    """Bound + gradient in IEEE binary128 (oracle/hp/sgpr_hp.c, -DUSE_QUAD): independent of float64 rounding altogether."""
    from oracle import hp
    X, y, Z, th = make_problem(N, M, D, seed=seed)
    F, g = hp.bound_grad(X.numpy(), y.numpy(), Z.numpy(), th.numpy(), jitter, "quad")
    np.savez(os.path.join(OUT, name + ".npz"), X=X.numpy(), y=y.numpy(), Z=Z.numpy(), theta=th.numpy(), jitter=jitter, F=F,
             d_ell=g["ell"], d_sf2=g["sf2"], d_s2=g["s2"], d_Z=g["Z"])
    print(name, F)


def svgp_case(name, N, M, D, B, seed):
    X, y, Z, th = make_problem(N, M, D, seed=seed)
    rs = np.random.RandomState(seed + 7)
    m = torch.tensor(0.3 * rs.randn(M))
    Ls = torch.tensor(np.tril(np.eye(M) + 0.1 * rs.randn(M, M)))
    xb, yb = X[:B], y[:B]
    ell, sf2, s2 = th[:D], th[D], th[D + 1]
    e_unw = svgp.svgp_elbo_unwhitened(xb, yb, Z, m, Ls, ell, sf2, s2, N, jitter=1e-6)
    yb01 = (yb > 0).double()
    e_bern = svgp.svgp_elbo(xb, yb01, Z, m, Ls, ell, sf2, s2, N, likelihood="bernoulli")
    np.savez(os.path.join(OUT, name + ".npz"), xb=xb.numpy(), yb=yb.numpy(), yb01=yb01.numpy(), Z=Z.numpy(), theta=th.numpy(),
             m=m.numpy(), Ls=Ls.numpy(), num_data=N, elbo_unwhitened=e_unw.item(), elbo_bernoulli_gh20=e_bern.item())
    print(name, e_unw.item(), e_bern.item())


def rng_streams():
    # R7: experiments/regression.py:203 -> utils/dataset.py:62-63 -> experiments/regression.py:83 on the GLOBAL legacy stream
    out = {}
    for split, N, prop, M in [(0, 9568, 0.8, 500), (3, 545, 0.9, 100), (7, 1000, 0.8, 20)]:
        np.random.seed(12345)            # args.seed: irrelevant, re-seeded below exactly as the reference does
        ind = np.arange(N)
        np.random.seed(173 + split)
        np.random.shuffle(ind)
        n = int(N * prop)
        z_idx = np.random.randint(0, n, M)
        out[f"split{split}_train_head"] = ind[:64].copy()
        out[f"split{split}_zidx"] = z_idx
    # R8: DataLoader(TensorDataset, batch_size, shuffle=True) under torch.manual_seed(seed)
    from torch.utils.data import TensorDataset, DataLoader
    for seed, n, bs in [(42, 7654, 1024), (7, 100, 32)]:
        torch.manual_seed(seed)
        ds = TensorDataset(torch.arange(n), torch.arange(n))
        dl = DataLoader(ds, batch_size=bs, shuffle=True)
        for ep in range(2):
            idx = torch.cat([xb for xb, _ in dl]).numpy()
            out[f"loader_seed{seed}_n{n}_bs{bs}_epoch{ep}"] = idx
    np.savez(os.path.join(OUT, "rng_streams.npz"), **out)
    print("rng", list(out)[:3])


if __name__ == "__main__":
    sgpr_case("sgpr_small_1d", 200, 12, 1, 11, 1e-6)
    sgpr_case("sgpr_small_3d", 300, 25, 3, 12, 1e-6)
    sgpr_case("sgpr_mid_4d", 600, 70, 4, 13, 1e-5)
    sgpr_hp_case("sgpr_hp_small", 500, 48, 3, 14, 1e-6)
    svgp_case("svgp_small", 400, 30, 3, 64, 21)
    rng_streams()
