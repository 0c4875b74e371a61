"""Stationary ARD kernels (oracle; torch CPU float64).

Reference call sites: ScaleKernel(RBFKernel(ard_num_dims=D)) at models/sgpr.py:36,
models/bayesian_sgpr_hmc.py:41, models/svgp.py:53; pymc3 ``sig_f**2 * ExpQuad(D, ls)`` at
models/bayesian_sgpr_hmc.py:65; Matern32 in experiments/co2_bayesian_sgpr_hmc.py:74-83.
"""
import math
import torch


def scaled_sqdist(X1, X2, ell):
    """d2[i,j] = sum_d ((X1[i,d]-X2[j,d])/ell[d])^2, evaluated directly (no ||a||^2-2ab+||b||^2)."""
    a = X1 / ell
    b = X2 / ell
    diff = a.unsqueeze(-2) - b.unsqueeze(-3)
    return (diff * diff).sum(-1)


def scaled_sqdist_gpytorch(X1, X2, ell):
    """gpytorch's evaluation order: centre on mean(X1), ||a||^2 - 2ab + ||b||^2, clamp>=0."""
    a = X1 / ell
    b = X2 / ell
    adjustment = a.mean(-2, keepdim=True)
    a = a - adjustment
    b = b - adjustment
    a2 = (a * a).sum(-1, keepdim=True)
    b2 = (b * b).sum(-1, keepdim=True)
    res = a2 - 2.0 * a @ b.transpose(-1, -2) + b2.transpose(-1, -2)
    return res.clamp_min(0.0)


def ard_kernel(X1, X2, ell, sf2, kind="rbf", gpytorch_order=False):
    """k(x,z) for kind in {rbf, matern32, matern52, ("rq", alpha)}; sf2 = outputscale (a variance).
    ("rq", alpha): rational quadratic (1 + d2 / (2 alpha))^(-alpha) -- gpytorch RQKernel / pymc3 RatQuad
    (experiments/co2_bayesian_sgpr_hmc.py:77,127), alpha a constant of the evaluation."""
    d2 = scaled_sqdist_gpytorch(X1, X2, ell) if gpytorch_order else scaled_sqdist(X1, X2, ell)
    if isinstance(kind, tuple):
        assert kind[0] == "rq"
        alpha = float(kind[1])
        return sf2 * torch.pow(1.0 + d2 / (2.0 * alpha), -alpha)
    if kind == "rbf":
        return sf2 * torch.exp(-0.5 * d2)
    # sqrt with a safe sub-gradient at 0 (d2 == 0 on the Kzz diagonal)
    r = torch.sqrt(d2.clamp_min(1e-300))
    r = torch.where(d2 > 0, r, torch.zeros_like(r))
    if kind == "matern32":
        a = math.sqrt(3.0)
        return sf2 * (1.0 + a * r) * torch.exp(-a * r)
    if kind == "matern52":
        a = math.sqrt(5.0)
        return sf2 * (1.0 + a * r + (5.0 / 3.0) * d2) * torch.exp(-a * r)
    raise ValueError(kind)


def kdiag(X, sf2):
    return sf2 * torch.ones(X.shape[0], dtype=X.dtype)
