"""Whitened SVGP minibatch ELBO and predictive (oracle; torch CPU float64).

Reference call sites: models/svgp.py:37-58 (CholeskyVariationalDistribution + whitened
VariationalStrategy), :90 VariationalELBO(num_data=N), :104-106 loss = -mll(self(x_b), y_b);
models/bayesian_svgp.py:129-133,157-167 (theta draws overwrite the kernel parameters);
scratch_pymc3.py:78 (Bernoulli likelihood).
"""
import math
import numpy as np
import torch

from .kernels import ard_kernel
from .linalg import psd_safe_cholesky

LOG2PI = math.log(2.0 * math.pi)
_GH_X, _GH_W = np.polynomial.hermite.hermgauss(20)
GH_X = torch.tensor(_GH_X, dtype=torch.float64)
GH_W = torch.tensor(_GH_W, dtype=torch.float64)

VAR_CHOL_JITTER_F64 = 1e-6   # gpytorch variational_cholesky_jitter, float64
DATA_DIAG_JITTER = 1e-4      # VariationalStrategy.forward: data_data_covar.add_jitter(1e-4)


def log_ndtr(x):
    return torch.special.log_ndtr(x)


def svgp_marginals(xb, Z, m, Ls_raw, ell, sf2, jitter_policy="gpytorch", kind="rbf"):
    """q(f_b) marginals of the whitened strategy: mu = a^T m ; var = k_bb + 1e-4 + ||L_s^T a||^2 - ||a||^2."""
    M = Z.shape[0]
    Ls = torch.tril(Ls_raw)
    Kzz = ard_kernel(Z, Z, ell, sf2, kind) + VAR_CHOL_JITTER_F64 * torch.eye(M, dtype=xb.dtype)
    with torch.no_grad():
        _, jit = psd_safe_cholesky(Kzz.detach(), jitter_policy)
    Lz = torch.linalg.cholesky(Kzz + jit * torch.eye(M, dtype=xb.dtype))
    Kzb = ard_kernel(Z, xb, ell, sf2, kind)
    a = torch.linalg.solve_triangular(Lz, Kzb, upper=False)
    mu = a.T @ m
    Lsa = Ls.T @ a
    var = sf2 + DATA_DIAG_JITTER + (Lsa * Lsa).sum(0) - (a * a).sum(0)
    return mu, var


def kl_whitened(m, Ls_raw):
    """KL(N(m, L_s L_s^T) || N(0, I)) = 1/2 [tr S + m^T m - M - 2 sum log|diag L_s|]."""
    Ls = torch.tril(Ls_raw)
    M = m.shape[0]
    return 0.5 * ((Ls * Ls).sum() + m @ m - M - 2.0 * torch.log(torch.abs(torch.diagonal(Ls))).sum())


def gaussian_expected_log_prob(yb, mu, var, s2):
    return -0.5 * (((yb - mu) ** 2 + var) / s2 + torch.log(s2) + LOG2PI)


def bernoulli_probit_expected_log_prob(yb, mu, var):
    """E_q[log Phi((2y-1) f)], 20-point Gauss-Hermite: (1/sqrt(pi)) sum_i w_i log Phi((2y-1)(mu+sqrt(2 var) x_i))."""
    sgn = 2.0 * yb - 1.0
    f = mu.unsqueeze(-1) + torch.sqrt(2.0 * var).unsqueeze(-1) * GH_X
    return (GH_W * log_ndtr(sgn.unsqueeze(-1) * f)).sum(-1) / math.sqrt(math.pi)


def svgp_elbo(xb, yb, Z, m, Ls_raw, ell, sf2, s2, num_data, likelihood="gaussian",
              jitter_policy="gpytorch", kind="rbf"):
    """VariationalELBO value: sum_b E_q[log p(y_b|f_b)] / B - KL / num_data   (models/svgp.py:90,106)."""
    mu, var = svgp_marginals(xb, Z, m, Ls_raw, ell, sf2, jitter_policy, kind)
    if likelihood == "gaussian":
        ell_b = gaussian_expected_log_prob(yb, mu, var, s2)
    elif likelihood == "bernoulli":
        ell_b = bernoulli_probit_expected_log_prob(yb, mu, var)
    else:
        raise ValueError(likelihood)
    return ell_b.sum() / xb.shape[0] - kl_whitened(m, Ls_raw) / num_data


def svgp_elbo_unwhitened(xb, yb, Z, m, Ls_raw, ell, sf2, s2, num_data, jitter):
    """Definition without whitening (known-answer anchor): q(u)=N(L_z m, L_z S L_z^T), p(u)=N(0,Kzz)."""
    M = Z.shape[0]
    Ls = torch.tril(Ls_raw)
    Kzz = ard_kernel(Z, Z, ell, sf2) + jitter * torch.eye(M, dtype=xb.dtype)
    Lz = torch.linalg.cholesky(Kzz)
    mu_u = Lz @ m
    S_u = Lz @ Ls @ Ls.T @ Lz.T
    Kinv = torch.linalg.inv(Kzz)
    Kbz = ard_kernel(xb, Z, ell, sf2)
    H = Kbz @ Kinv
    mu = H @ mu_u
    var = sf2 + DATA_DIAG_JITTER - (H * Kbz).sum(-1) + ((H @ S_u) * H).sum(-1)
    q = torch.distributions.MultivariateNormal(mu_u, covariance_matrix=0.5 * (S_u + S_u.T))
    p = torch.distributions.MultivariateNormal(torch.zeros(M, dtype=xb.dtype), covariance_matrix=Kzz)
    kl = torch.distributions.kl_divergence(q, p)
    return gaussian_expected_log_prob(yb, mu, var, s2).sum() / xb.shape[0] - kl / num_data


def svgp_predict(xs, Z, m, Ls_raw, ell, sf2, s2, jitter_policy="gpytorch", add_noise=True):
    """posterior_predictive of models/svgp.py:132-141, diagonal only (the exact expression; SURVEY A.7 note on
    fast_pred_var)."""
    mu, var = svgp_marginals(xs, Z, m, Ls_raw, ell, sf2, jitter_policy)
    return mu, var + (s2 if add_noise else 0.0)
