"""Composite covariance functions (oracle; torch CPU float64) -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline).

The reference's CO2 model is a sum of scaled products of stationary kernels:
  gpytorch  experiments/co2_bayesian_sgpr_hmc.py:74-83   ScaleKernel(Periodic * RBF) + ScaleKernel(RBF) + ScaleKernel(RQ) + ScaleKernel(RBF)
  pymc3     experiments/co2_bayesian_sgpr_hmc.py:107-149 n_per^2 Periodic(1, ls) ExpQuad(l_pdecay) + n_med^2 RatQuad(l_med, alpha)
                                                         + n_trend^2 ExpQuad(l_trend) + n_noise^2 Matern32(l_noise), noise sigma
gpytorch and pymc3 are not part of the reference tree; the kernel formulas restated here are their published ones, in ONE canonical
parameterisation (the one include/ggp_b200.h documents), with the maps from the two libraries' parameters next to the targets:
  rbf       exp(-d2/2), d2 = sum_c ((x_c - z_c)/ell_c)^2       matern32 / matern52  as oracle/kernels.py
  rq        (1 + d2/(2 alpha))^(-alpha)
  periodic  exp(-2 sum_c sin^2(pi (x_c - z_c)/p_c) / ell_c^2)   (pymc3 Periodic(ls): ell = 2 ls; gpytorch PeriodicKernel: ell^2 = lengthscale)
Parameter row of a program, in program order: for each term a_t, then for each factor ell[d], then rq: alpha | periodic: period[d].
Everything is differentiable torch: gradients of the bound come from autograd (what the reference's loss.backward() / theano do).
"""
import math

import torch

from .linalg import psd_safe_cholesky
from .sgpr import LOG2PI, _tri_solve

NPAR = {"rbf": lambda d: d, "matern32": lambda d: d, "matern52": lambda d: d, "rq": lambda d: d + 1, "periodic": lambda d: 2 * d}


def nparams(prog, d):
    return sum(1 + sum(NPAR[f](d) for f in term) for term in prog)


def amplitude_indices(prog, d):
    idx, p = [], 0
    for term in prog:
        idx.append(p)
        p += 1 + sum(NPAR[f](d) for f in term)
    return idx


def _factor(kind, p, X1, X2, d):
    ell = p[:d]
    diff = X1.unsqueeze(-2) - X2.unsqueeze(-3)                # [n1, n2, d]
    if kind == "periodic":
        per = p[d:2 * d]
        s = torch.sin(math.pi * diff / per) / ell
        return torch.exp(-2.0 * (s * s).sum(-1))
    t = diff / ell
    d2 = (t * t).sum(-1)
    if kind == "rbf":
        return torch.exp(-0.5 * d2)
    if kind == "rq":
        alpha = p[d]
        return torch.exp(-alpha * torch.log1p(d2 / (2.0 * alpha)))
    r = torch.sqrt(d2.clamp_min(1e-300))
    r = torch.where(d2 > 0, r, torch.zeros_like(r))
    if kind == "matern32":
        a = math.sqrt(3.0)
        return (1.0 + a * r) * torch.exp(-a * r)
    if kind == "matern52":
        a = math.sqrt(5.0)
        return (1.0 + a * r + (5.0 / 3.0) * d2) * torch.exp(-a * r)
    raise ValueError(kind)


def composite_kernel(prog, kth, X1, X2):
    """k(X1, X2)[n1, n2] for the program `prog` (tuple of terms, each a tuple of factor names) at the parameter row kth[P]."""
    d = X1.shape[-1]
    K, p = 0.0, 0
    for term in prog:
        v = kth[p]
        p += 1
        for f in term:
            n = NPAR[f](d)
            v = v * _factor(f, kth[p:p + n], X1, X2, d)
            p += n
        K = K + v
    return K


def composite_kdiag(prog, kth, d):
    return sum(kth[i] for i in amplitude_indices(prog, d))


def sgpr_bound_composite(X, y, Z, prog, kth, s2, jitter_policy="gpytorch", normalize="n", return_state=False):
    """oracle.sgpr.sgpr_bound with the composite kernel (same A/B form, same jitter rule)."""
    N, M, d = X.shape[0], Z.shape[0], X.shape[1]
    Kzz = composite_kernel(prog, kth, Z, Z)
    with torch.no_grad():
        _, jit = psd_safe_cholesky(Kzz.detach(), jitter_policy)
    L = torch.linalg.cholesky(Kzz + jit * torch.eye(M, dtype=X.dtype))
    Kzx = composite_kernel(prog, kth, Z, X)
    A = _tri_solve(L, Kzx)
    S = A @ A.T
    B = torch.eye(M, dtype=X.dtype) + S / s2
    LB = torch.linalg.cholesky(B)
    b = A @ y
    c = _tri_solve(LB, b.unsqueeze(-1)).squeeze(-1) / s2
    kd = composite_kdiag(prog, kth, d)
    F = (-0.5 * N * LOG2PI - 0.5 * N * torch.log(s2) - torch.log(torch.diagonal(LB)).sum()
         - 0.5 * ((y @ y) / s2 - c @ c) - 0.5 * (N * kd - torch.trace(S)) / s2)
    out = F / N if normalize == "n" else F
    if return_state:
        return out, dict(L=L, LB=LB, A=A, b=b, c=c, jitter=jit)
    return out


def sgpr_bound_and_grads_composite(X, y, Z, prog, kth, s2, jitter_policy="gpytorch", normalize="n"):
    kth = kth.detach().clone().requires_grad_(True)
    s2 = s2.detach().clone().requires_grad_(True)
    Z = Z.detach().clone().requires_grad_(True)
    F = sgpr_bound_composite(X, y, Z, prog, kth, s2, jitter_policy, normalize)
    g = torch.autograd.grad(F, [kth, s2, Z])
    return F.detach(), dict(k=g[0], s2=g[1], Z=g[2])


def sgpr_predict_composite(Xs, X, y, Z, prog, kth, s2, jitter_policy="gpytorch", train_diag_correction=True):
    """Eval-mode sparse predictive (oracle.sgpr.sgpr_predict) with the composite kernel: mean, full covariance incl. likelihood noise."""
    N, M, d = X.shape[0], Z.shape[0], X.shape[1]
    Kzz = composite_kernel(prog, kth, Z, Z)
    L, jit = psd_safe_cholesky(Kzz, jitter_policy)
    A = _tri_solve(L, composite_kernel(prog, kth, Z, X))
    kd = composite_kdiag(prog, kth, d)
    lam = s2 + ((kd - (A * A).sum(0)).clamp_min(0.0) if train_diag_correction else torch.zeros(N, dtype=X.dtype))
    Aw = A / lam
    B = torch.eye(M, dtype=X.dtype) + Aw @ A.T
    LB = torch.linalg.cholesky(B)
    c = _tri_solve(LB, (Aw @ y).unsqueeze(-1)).squeeze(-1)
    a = _tri_solve(L, composite_kernel(prog, kth, Z, Xs))
    t = _tri_solve(LB, a)
    mean = t.T @ c
    cov = t.T @ t + torch.diag((kd - (a * a).sum(0)).clamp_min(0.0)) + s2 * torch.eye(Xs.shape[0], dtype=X.dtype)
    return mean, cov


# ---- the pymc3 CO2 target (experiments/co2_bayesian_sgpr_hmc.py:107-152) ---------------------------------------------------------
CO2_PROG = (("periodic", "rbf"), ("rq",), ("rbf",), ("matern32",))
# unconstrained point x[11] = (log_n_per, log_l_pdecay, log_l_psmooth, log_n_med, log_l_med, log_alpha, log_n_trend, log_l_trend,
#                             log_n_noise, log_l_noise, sigma_log__)
CO2_PRIOR_SD = (3.0, 0.1, 1.0, 3.0, 3.0, 0.1, 3.0, 1.0, 3.0, 1.0)


def co2_params_from_x(x):
    """(kth[11], s2) of CO2_PROG in D = 1 from the pymc3 point: Periodic(1, period=1, ls=l_psmooth) is the canonical periodic factor
    with ell = 2 l_psmooth, period 1; amplitudes are n^2; noise variance sigma^2."""
    e = torch.exp(x)
    one = torch.ones((), dtype=x.dtype)
    kth = torch.stack([e[0] ** 2, 2.0 * e[2], one, e[1],      # seasonal: a, periodic ell, period, rbf ell
                       e[3] ** 2, e[4], e[5],                 # medium: a, ell, alpha
                       e[6] ** 2, e[7],                       # trend
                       e[8] ** 2, e[9]])                      # noise (Matern32)
    return kth, e[10] ** 2


def co2_logp(x, X, y, Z, jitter_policy="pymc3"):
    kth, s2 = co2_params_from_x(x)
    F = sgpr_bound_composite(X, y, Z, CO2_PROG, kth, s2, jitter_policy, normalize="none")
    sd = torch.tensor(CO2_PRIOR_SD, dtype=x.dtype)
    lp_normal = (-0.5 * (x[:10] / sd) ** 2 - torch.log(sd) - 0.5 * LOG2PI).sum()
    sigma = torch.exp(x[10])
    lp_sigma = 0.5 * math.log(2.0 / math.pi) - 0.5 * sigma ** 2 + x[10]     # HalfNormal(1) + log-transform Jacobian
    return F + lp_normal + lp_sigma


def co2_logp_dlogp(x, X, y, Z, jitter_policy="pymc3"):
    x = x.detach().clone().requires_grad_(True)
    lp = co2_logp(x, X, y, Z, jitter_policy)
    (g,) = torch.autograd.grad(lp, x)
    return lp.detach(), g
