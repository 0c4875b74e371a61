"""gpflow SGPMC log-posterior density that tfp HMC samples in models/sgp_hmc.py:38-83 (oracle),
and the Bernoulli-probit variant BASELINE.json config 5 names.

State (unconstrained): v [M] (whitened inducing values, prior N(0,I)), raw lengthscales [D], raw kernel
variance, raw noise variance.  Positive parameters use softplus; the priors Gamma(2,1) are on the constrained
values (models/sgp_hmc.py:47-49) and gpflow adds log|d constrained / d unconstrained| = log sigmoid(raw).
log_posterior_density = sum_n var_exp(mu_n, var_n, y_n) + log N(v;0,I) + sum log-prior(+Jacobian)
with mu = A^T v, var = k_nn - colsum(A^2), A = L^{-1} Kzx, L = chol(Kzz + jitter I), jitter 1e-5 (models/sgp_hmc.py:20).
"""
import math
import torch
import torch.nn.functional as Fnn

from .kernels import ard_kernel
from .svgp import gaussian_expected_log_prob, bernoulli_probit_expected_log_prob

LOG2PI = math.log(2.0 * math.pi)


def sgpmc_logp(v, raw, X, y, Z, likelihood="gaussian", jitter=1e-5, with_priors=True):
    """raw = [raw_ell[D], raw_sf2, raw_s2] (softplus-unconstrained; raw_s2 ignored for bernoulli)."""
    D = X.shape[1]
    M = Z.shape[0]
    ell = Fnn.softplus(raw[:D])
    sf2 = Fnn.softplus(raw[D])
    Kzz = ard_kernel(Z, Z, ell, sf2) + jitter * torch.eye(M, dtype=X.dtype)
    L = torch.linalg.cholesky(Kzz)
    A = torch.linalg.solve_triangular(L, ard_kernel(Z, X, ell, sf2), upper=False)
    mu = A.T @ v
    var = sf2 - (A * A).sum(0)
    if likelihood == "gaussian":
        s2 = Fnn.softplus(raw[D + 1])
        ll = gaussian_expected_log_prob(y, mu, var, s2).sum()
        npos = D + 2
    else:
        ll = bernoulli_probit_expected_log_prob(y, mu, var).sum()
        npos = D + 1
    lp = ll - 0.5 * (v @ v) - 0.5 * M * LOG2PI
    if with_priors:
        pos = Fnn.softplus(raw[:npos])
        lp = lp + (torch.log(pos) - pos).sum() + Fnn.logsigmoid(raw[:npos]).sum()
    return lp


def sgpmc_logp_dlogp(v, raw, X, y, Z, likelihood="gaussian", jitter=1e-5, with_priors=True):
    v = v.detach().clone().requires_grad_(True)
    raw = raw.detach().clone().requires_grad_(True)
    lp = sgpmc_logp(v, raw, X, y, Z, likelihood, jitter, with_priors)
    gv, gr = torch.autograd.grad(lp, [v, raw], allow_unused=True)
    if gr is None:
        gr = torch.zeros_like(raw)
    return lp.detach(), gv, gr


def sgpmc_logp_dlogp_chunked(v, raw, X, y, Z, likelihood="gaussian", jitter=1e-5, with_priors=True, chunk=8192):
    """Same value and gradient, the data term accumulated over row chunks (the log-likelihood is a sum over n; each chunk runs its
    own autograd pass through chol(Kzz)), so that BASELINE configs[4] (N = 2e5, D = 16, M = 512) fits in host memory: the unchunked
    evaluation materialises an [M, N, D] difference tensor (13 GB) and keeps it for the backward pass."""
    N, M = X.shape[0], Z.shape[0]
    lp = torch.zeros((), dtype=X.dtype)
    gv, gr = torch.zeros_like(v), torch.zeros_like(raw)
    for i0 in range(0, N, chunk):
        l, a, b = sgpmc_logp_dlogp(v, raw, X[i0:i0 + chunk], y[i0:i0 + chunk], Z, likelihood, jitter, with_priors=False)
        # every chunk call adds log N(v; 0, I) once: keep it only once
        extra = -0.5 * (v @ v) - 0.5 * M * LOG2PI
        lp = lp + l - extra
        gv = gv + a + v
        gr = gr + b
    vv = v.detach().clone().requires_grad_(True)
    rr = raw.detach().clone().requires_grad_(True)
    D = X.shape[1]
    pr = -0.5 * (vv @ vv) - 0.5 * M * LOG2PI
    if with_priors:
        npos = D + 2 if likelihood == "gaussian" else D + 1
        pos = Fnn.softplus(rr[:npos])
        pr = pr + (torch.log(pos) - pos).sum() + Fnn.logsigmoid(rr[:npos]).sum()
    a, b = torch.autograd.grad(pr, [vv, rr], allow_unused=True)
    return lp + pr.detach(), gv + a, gr + (b if b is not None else torch.zeros_like(raw))
