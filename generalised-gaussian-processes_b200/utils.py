"""Consumers of the predictive mean / variance (host code; same names and argument meaning as the reference).

  get_posterior_predictive_means_stds, get_posterior_predictive_mean, get_posterior_predictive_uncertainty_intervals
        utils/posterior_predictive.py:12-46
  rmse, nlpd, nlpd_marginal, nlpd_mixture                                     utils/metrics.py:38-67
"""
import math

import numpy as np
import torch


def get_posterior_predictive_means_stds(Y_test_pred_list):
    sample_means = torch.stack([d.loc.detach() for d in Y_test_pred_list])
    sample_stds = torch.stack([d.variance.detach().sqrt() for d in Y_test_pred_list])
    return sample_means, sample_stds


def get_posterior_predictive_mean(sample_means):
    return torch.mean(sample_means, axis=0)


def get_posterior_predictive_uncertainty_intervals(sample_means, sample_stds, n_draws=1000, generator=None):
    """95 % interval of the equally weighted Gaussian mixture at every test point.  The reference loops over test points and
    samples 1000 mixture draws each (utils/posterior_predictive.py:37-45); this draws all points at once on the device."""
    C, T = sample_means.shape
    comp = torch.randint(0, C, (n_draws, T), device=sample_means.device, generator=generator)
    eps = torch.randn(n_draws, T, dtype=sample_means.dtype, device=sample_means.device, generator=generator)
    cols = torch.arange(T, device=sample_means.device).expand(n_draws, T)
    draws = sample_means[comp, cols] + sample_stds[comp, cols] * eps
    q = torch.quantile(draws, torch.tensor([0.025, 0.975], dtype=draws.dtype, device=draws.device), dim=0)
    return q[0].cpu().numpy(), q[1].cpu().numpy()


def rmse(Y_pred_mean, Y_test, Y_std):
    return float(Y_std) * torch.sqrt(torch.mean((Y_pred_mean - Y_test) ** 2)).detach()


def nlpd(Y_test_pred, Y_test, Y_std):
    """-(joint MVN log-density / N* - log Y_std): uses the FULL predictive covariance like utils/metrics.py:42-47."""
    lpd = Y_test_pred.log_prob(Y_test)
    return -(lpd.detach() / len(Y_test) - math.log(float(Y_std)))


def nlpd_marginal(Y_test_pred, Y_test, Y_std):
    var = Y_test_pred.variance.detach()
    lp = -0.5 * ((Y_test - Y_test_pred.loc.detach()) ** 2 / var + torch.log(var) + math.log(2 * math.pi)) - math.log(float(Y_std))
    return float(-lp.mean())


def nlpd_mixture(Y_test_pred_list, Y_test, Y_std):
    return float(np.mean([float(nlpd(p, Y_test, Y_std)) for p in Y_test_pred_list]))
