"""CPU tests of the host-side model pieces that need no kernel: q(log theta) of BayesianSVGP against a literal torch restatement of
the reference (models/bayesian_svgp.py:51-84, 129-133), the pymc3-style running mass matrix, and the opt-in row-shard reduction."""
import math

import pytest
import torch

import ggp_b200.models as mdl
from ggp_b200 import hmc as H
from ggp_b200.engine import Engine


def _reference_construct_sigma(q_sigma_vec, hyper_dim):
    """models/bayesian_svgp.py:51-61 restated: start from the IDENTITY, overwrite every lower-triangular entry (diagonal included)
    from q_sigma_vec in tril_indices order with a Python loop, Sigma = L L^T + 1e-5 I."""
    row_ids, col_ids = torch.tril_indices(hyper_dim, hyper_dim)
    lower = torch.eye(hyper_dim, dtype=torch.float64)
    k = 0
    for i, j in zip(row_ids, col_ids):
        lower[i, j] = q_sigma_vec[k]
        k += 1
    return lower @ lower.T + torch.eye(hyper_dim, dtype=torch.float64) * 1e-5


def test_variational_hyper_dist_sigma_kl_and_theta_map():
    torch.manual_seed(3)
    D, n = 3, 777
    q = mdl.VariationalHyperDist(D + 2, n)
    with torch.no_grad():
        q.q_mu.copy_(torch.randn(D + 2, dtype=torch.float64) * 0.3)
        q.q_sigma_vec.copy_(torch.randn((D + 2) * (D + 3) // 2, dtype=torch.float64) * 0.4)
    sig_ref = _reference_construct_sigma(q.q_sigma_vec.detach(), D + 2)
    assert torch.allclose(q.construct_sigma(), sig_ref, rtol=0, atol=1e-15)
    # KL(q || N(0, 0.01 I)) / n (models/bayesian_svgp.py:81-84, prior :112-113) against the closed form
    mu, hd = q.q_mu.detach(), D + 2
    kl = 0.5 * (torch.trace(sig_ref) / 0.01 + (mu @ mu) / 0.01 - hd + hd * math.log(0.01) - torch.logdet(sig_ref))
    assert abs(float(q.kl_per_point()) - float(kl) / n) < 1e-12 * abs(float(kl) / n)
    ref = torch.distributions.kl_divergence(torch.distributions.MultivariateNormal(mu, sig_ref),
                                            torch.distributions.MultivariateNormal(torch.zeros(hd, dtype=torch.float64),
                                                                                   0.01 * torch.eye(hd, dtype=torch.float64))) / n
    assert abs(float(q.kl_per_point()) - float(ref)) < 1e-12 * abs(float(ref))
    # theta map (models/bayesian_svgp.py:129-133): theta[0] -> outputscale, theta[1:-1] -> lengthscale, theta[-1]^2 -> noise;
    # engine rows are [ell[D], sf2, s2]
    draws = q(4).detach()
    X = torch.zeros(10, D, dtype=torch.float64)
    m = mdl.BayesianStochasticVariationalGP(X, torch.zeros(10, dtype=torch.float64), mdl.GaussianLikelihood(), torch.zeros(4, D, dtype=torch.float64))
    th = m._thetas(draws)
    e = torch.exp(draws)
    assert torch.equal(th[:, :D], e[:, 1:-1]) and torch.equal(th[:, D], e[:, 0]) and torch.equal(th[:, D + 1], e[:, -1] ** 2)
    # the KL term carries the gradient of q (upstream the property setters detach theta from the data term)
    q.kl_per_point().backward()
    assert q.q_mu.grad is not None and q.q_sigma_vec.grad is not None and q.q_sigma_vec.grad.abs().sum() > 0


def test_running_diag_mass_is_pymc3s_weighted_variance():
    """QuadPotentialDiagAdapt: foreground starts at (mean x0, variance 1, weight 10), Welford updates, variance = raw / n_samples."""
    g = torch.Generator().manual_seed(0)
    x0 = torch.tensor([[0.5, -1.0]], dtype=torch.float64)
    r = H.RunningDiagMass(x0, window=5)
    xs = [x0 + 0.3 * torch.randn(1, 2, dtype=torch.float64, generator=g) for _ in range(7)]
    # literal restatement of pymc3's _WeightedVariance
    n, mean, raw = 10, x0.clone(), torch.ones_like(x0) * 10
    bn, bmean, braw = 0, torch.zeros_like(x0), torch.zeros_like(x0)
    for i, x in enumerate(xs):
        n += 1; od = x - mean; mean = mean + od / n; raw = raw + od * (x - mean)
        bn += 1; od = x - bmean; bmean = bmean + od / bn; braw = braw + od * (x - bmean)
        var = r.update(x)
        assert torch.allclose(var, raw / n, rtol=0, atol=1e-15)
        if (i + 1) % 5 == 0:
            n, mean, raw = bn, bmean, braw
            bn, bmean, braw = 0, torch.zeros_like(x0), torch.zeros_like(x0)


def test_nuts_with_short_tuning_adapts_the_mass_matrix():
    """tune = 25 (models/bayesian_sgpr_hmc.py:145): the running estimator moves the metric from the first draw on."""
    g = torch.Generator().manual_seed(1)
    prec = torch.diag(torch.tensor([100.0, 0.04], dtype=torch.float64))
    f = lambda x: (-0.5 * ((x @ prec) * x).sum(1), -(x @ prec))
    res = H.nuts_sample(f, torch.zeros(4, 2, dtype=torch.float64), 10, tune=25, generator=g)
    im = res["inv_mass"]
    assert (im[:, 0] < 0.95).all() and not torch.allclose(im, torch.ones_like(im))
    assert torch.isfinite(res["samples"]).all()


def test_row_shard_reduction_is_opt_in():
    assert Engine._resolve_group(None) == (False, None) and Engine._resolve_group(False) == (False, None)
    with pytest.raises(RuntimeError):
        Engine._resolve_group(True)     # no process group initialised
    sentinel = object()
    assert Engine._resolve_group(sentinel) == (True, sentinel)
