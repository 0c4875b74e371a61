"""Synthetic generators for the five BASELINE.json configurations and the reference's index streams.

All generators are float64, seeded with numpy.random.RandomState(173 + config index) mirroring BASE_SEED
(utils/config.py:14); X is standardised like utils/dataset.py:38-41.  There is no network, so the UCI files the
reference downloads are replaced by data of the same shape.
"""
import numpy as np
import torch

BASE_SEED = 173  # utils/config.py:14


def _standardise(a):
    return (a - a.mean(0)) / (1e-6 + a.std(0))  # utils/dataset.py:38-41


def select_inducing_indices(n_train, num_inducing, rng=None):
    """Z_init index draw of experiments/regression.py:83: np.random.randint(0, len(X_train), M) -- WITH replacement,
    on the legacy MT19937 stream.  `rng` is a numpy RandomState (or the global numpy module state if None)."""
    r = np.random if rng is None else rng
    return r.randint(0, n_train, num_inducing)


def dataset_split_stream(N, split, prop=0.8):
    """Index stream of utils/dataset.py:60-66: seed BASE_SEED+split, shuffle arange(N), first floor(prop*N) train.
    Returns (rng, train_idx, test_idx); `rng` continues the same legacy stream (the Z draw comes next on it)."""
    rng = np.random.RandomState(BASE_SEED + split)
    ind = np.arange(N)
    rng.shuffle(ind)
    n = int(N * prop)
    return rng, ind[:n], ind[n:]


def minibatch_indices(n, batch_size, generator=None):
    """Per-epoch minibatch index lists of DataLoader(TensorDataset(X, Y), batch_size, shuffle=True)
    (experiments/regression.py:107-108; torch RandomSampler semantics): one int64 seed is drawn from
    `generator` (global torch CPU generator if None) with .random_(), a fresh generator is seeded with it,
    randperm(n) is cut into consecutive slices, the last partial batch is kept."""
    seed = int(torch.empty((), dtype=torch.int64).random_(generator=generator).item())
    g = torch.Generator()
    g.manual_seed(seed)
    perm = torch.randperm(n, generator=g)
    return [perm[i:i + batch_size] for i in range(0, n, batch_size)]


def _regression_xy(rs, N, D, noise=0.1):
    X = _standardise(rs.randn(N, D))
    w1, w2 = rs.randn(D), rs.randn(D)
    y = np.sin(X @ w1) + 0.5 * (X @ w2) + noise * rs.randn(N)
    y = (y - y.mean()) / y.std()
    return X, y


def config1_demo_1d(N=1000, M=20):
    """demo_1d_regression: models/sgpr.py:19-20,168-173."""
    rs = np.random.RandomState(BASE_SEED + 1)
    x = rs.randn(N) * 2 - 1
    y = np.sin(x * 3) + 0.3 * np.cos(x * 4 * 3.14) + 0.2 * rs.randn(N)
    Z = rs.randn(M)
    return dict(X=x[:, None], y=y, Z=Z[:, None], name="demo_1d_regression")


def config2_co2_shaped(N=545, M=100):
    """co2-shaped series (shape of experiments/co2_bayesian_sgpr_hmc.py:24-52)."""
    rs = np.random.RandomState(BASE_SEED + 2)
    t = np.linspace(0, 45.4, N)
    y = 0.03 * t ** 2 + np.sin(2 * np.pi * t) + 0.05 * rs.randn(N)
    y = (y - y.mean()) / y.std()
    X = _standardise(t[:, None])
    idx = select_inducing_indices(N, M, rs)
    return dict(X=X, y=y, Z=X[idx].copy(), Z_idx=idx, name="co2_shaped")


def config3_power_shaped(N=9568, D=4, M=500, split=0, prop=0.8):
    """UCI-Power-shaped regression (utils/dataset.py:185-196), split and Z draw by the reference's stream."""
    rs = np.random.RandomState(BASE_SEED + 3)
    X, y = _regression_xy(rs, N, D)
    rng, tr, te = dataset_split_stream(N, split, prop)
    idx = select_inducing_indices(len(tr), M, rng)
    Xtr = X[tr]
    return dict(X=Xtr, y=y[tr], X_test=X[te], y_test=y[te], Z=Xtr[idx].copy(), Z_idx=idx, name="power_shaped")


def config4_large(N=1_000_000, D=8, M=1024, with_replacement=False):
    """large_scale_regression-shaped synthetic, N=1e6, D=8, M=1024 (BASELINE.json configs[3])."""
    rs = np.random.RandomState(BASE_SEED + 4)
    X, y = _regression_xy(rs, N, D)
    idx = select_inducing_indices(N, M, rs) if with_replacement else rs.permutation(N)[:M]
    return dict(X=X, y=y, Z=X[idx].copy(), Z_idx=idx, name="large_scale_regression")


def theta_init_gpytorch(D):
    """gpytorch initial values (all raw parameters 0): ell = sf2 = ln 2, s2 = ln 2 + 1e-4."""
    return np.concatenate([np.full(D, np.log(2.0)), [np.log(2.0), np.log(2.0) + 1e-4]])


def theta_trained_like(D):
    return np.concatenate([np.full(D, np.sqrt(D)), [1.0, 0.1]])


def config5_classification(N=200_000, D=16, M=512):
    """Bernoulli-probit classification, N=2e5, D=16, M=512 (BASELINE.json configs[4])."""
    from scipy.special import ndtr
    rs = np.random.RandomState(BASE_SEED + 5)
    X = _standardise(rs.randn(N, D))
    w = rs.randn(D) / np.sqrt(D)
    f = 2.0 * np.sin(X @ w) + X[:, 0] * X[:, 1]
    y = (ndtr(f) > rs.rand(N)).astype(np.float64)
    idx = rs.permutation(N)[:M]
    return dict(X=X, y=y, Z=X[idx].copy(), Z_idx=idx, name="bernoulli_classification")
