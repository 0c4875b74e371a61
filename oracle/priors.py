"""pymc3 log-posterior that NUTS differentiates in models/bayesian_sgpr_hmc.py:60-78 (oracle).

Free variables, in declaration order, in pymc3's automatic log-transformed space:
  ls_log__[D], sig_f_log__, sig_n_log__
logp(x) = F(ell=e^x_ls, sf2=(e^x_f)^2, s2=(e^x_n)^2)             MarginalSparse VFE, jitter 1e-6, not / N
        + sum_d Gamma(alpha=2,beta=1).logp(ell_d)                  models/bayesian_sgpr_hmc.py:62
        + HalfCauchy(1).logp(sig_f) + HalfCauchy(1).logp(sig_n)    models/bayesian_sgpr_hmc.py:63,68
        + sum(x)                                                   log-Jacobians of the D+2 exp transforms
"""
import math
import torch

from .sgpr import sgpr_bound

LOG2 = math.log(2.0)
LOGPI = math.log(math.pi)


def gamma21_logp(v):
    # alpha*log(beta) - lgamma(alpha) + (alpha-1) log v - beta v, alpha=2, beta=1
    return torch.log(v) - v


def halfcauchy1_logp(v):
    return LOG2 - LOGPI - torch.log1p(v * v)


def unpack_theta(x, D):
    """Unconstrained x[D+2] -> (ell[D], sf2, s2) per update_model_to_hyper (models/bayesian_sgpr_hmc.py:82-86)."""
    ell = torch.exp(x[:D])
    sig_f = torch.exp(x[D])
    sig_n = torch.exp(x[D + 1])
    return ell, sig_f ** 2, sig_n ** 2


def log_prior_and_jacobian(x, D):
    ell = torch.exp(x[:D])
    sig_f = torch.exp(x[D])
    sig_n = torch.exp(x[D + 1])
    return gamma21_logp(ell).sum() + halfcauchy1_logp(sig_f) + halfcauchy1_logp(sig_n) + x.sum()


def sgpr_vfe_logp(x, X, y, Z, jitter_policy="pymc3"):
    D = X.shape[1]
    ell, sf2, s2 = unpack_theta(x, D)
    F = sgpr_bound(X, y, Z, ell, sf2, s2, jitter_policy=jitter_policy, normalize="none")
    return F + log_prior_and_jacobian(x, D)


def sgpr_vfe_logp_dlogp(x, X, y, Z, jitter_policy="pymc3"):
    x = x.detach().clone().requires_grad_(True)
    lp = sgpr_vfe_logp(x, X, y, Z, jitter_policy)
    (g,) = torch.autograd.grad(lp, x)
    return lp.detach(), g


def all_in_hmc_logp(x, X, y, M, jitter_policy="pymc3"):
    """models/all_in_HMC.py:47-60: theta AND the inducing inputs are sampled.  Free variables in declaration order:
    ls_log__[D], sig_f_log__, sig_n_log__, Z[M, D] (untransformed, Normal(0, 1) elementwise, :57).  x = [D + 2 + M*D]."""
    D = X.shape[1]
    Z = x[D + 2:].reshape(M, D)
    lpz = (-0.5 * Z * Z - 0.5 * math.log(2.0 * math.pi)).sum()
    return sgpr_vfe_logp(x[:D + 2], X, y, Z, jitter_policy) + lpz


def all_in_hmc_logp_dlogp(x, X, y, M, jitter_policy="pymc3"):
    x = x.detach().clone().requires_grad_(True)
    lp = all_in_hmc_logp(x, X, y, M, jitter_policy)
    (g,) = torch.autograd.grad(lp, x)
    return lp.detach(), g
