"""CPU tests of the host-side samplers (hmc.py is pure torch): they target analytic densities, no GPU and no oracle needed."""
import math

import torch

import ggp_b200.hmc as H


def _gauss_target(mu, prec):
    def f(x):
        d = x - mu
        return -0.5 * ((d @ prec) * d).sum(1), -(d @ prec)
    return f


def test_ckpt_ranges_of_the_iterative_uturn_scheme():
    # leaf 7 closes the subtrees [6,7], [4..7], [0..7] whose first leaves 6, 4, 0 were saved in slots 2, 1, 0
    assert H._ckpt_range(7) == (0, 2)
    assert H._ckpt_range(5) == (1, 1) and H._ckpt_range(3) == (0, 1) and H._ckpt_range(1) == (0, 0)
    assert H._ckpt_range(6)[1] == 2 and H._ckpt_range(4)[1] == 1 and H._ckpt_range(0)[1] == 0   # save slots of even leaves
    assert H._ckpt_range(13) == (2, 2)     # 13 = 0b1101 closes only [12, 13]; 12 = 0b1100 was saved in slot popcount(6) = 2


def test_nuts_recovers_a_correlated_gaussian():
    g = torch.Generator().manual_seed(0)
    mu = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    A = torch.tensor([[1.0, 0.6, 0.0], [0.6, 2.0, -0.3], [0.0, -0.3, 0.5]], dtype=torch.float64)
    f = _gauss_target(mu, torch.linalg.inv(A))
    x0 = torch.zeros(8, 3, dtype=torch.float64)
    res = H.nuts_sample(f, x0, 400, tune=300, generator=g)
    s = res["samples"].reshape(-1, 3)
    assert (s.mean(0) - mu).abs().max() < 0.15
    assert (torch.cov(s.T) - A).abs().max() < 0.25
    assert 0.6 < float(res["accept_rate"].mean()) < 0.97
    assert int(res["tree_depth"].max()) <= 10 and int(res["tree_depth"].min()) >= 1
    assert not bool(res["diverging"].any())
    # a tree of depth j holds 2^j - 1 new leaves at most (the last doubling may stop early)
    assert bool((res["n_leapfrog"] <= 2 ** res["tree_depth"].clamp(min=1) * 2).all())


def test_fixed_length_hmc_recovers_a_gaussian():
    g = torch.Generator().manual_seed(1)
    mu = torch.tensor([0.5, -1.0], dtype=torch.float64)
    prec = torch.diag(torch.tensor([4.0, 0.25], dtype=torch.float64))
    res = H.hmc_sample(_gauss_target(mu, prec), torch.zeros(6, 2, dtype=torch.float64), 500, tune=300, n_leapfrog=8,
                       step_size=0.1, generator=g)
    s = res["samples"].reshape(-1, 2)
    assert (s.mean(0) - mu).abs().max() < 0.2
    assert abs(float(s[:, 0].var()) - 0.25) < 0.08 and abs(float(s[:, 1].var()) - 4.0) < 1.2
