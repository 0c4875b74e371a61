"""Seeded synthetic problems shared by the oracle and GPU parity tests."""
import numpy as np
import torch


def make_problem(N, M, D, seed=0, without_replacement=True, noise=0.1):
    rs = np.random.RandomState(seed)
    X = rs.randn(N, D)
    w1, w2 = rs.randn(D), rs.randn(D)
    y = np.sin(X @ w1) + 0.5 * (X @ w2) + noise * rs.randn(N)
    y = (y - y.mean()) / y.std()
    idx = rs.permutation(N)[:M] if without_replacement else rs.randint(0, N, M)
    Z = X[idx].copy()
    ell = 0.8 + 0.6 * rs.rand(D) * np.sqrt(D)
    theta = np.concatenate([ell, [1.3, 0.2]])
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64)
    return t(X), t(y), t(Z), t(theta)


def relerr(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
