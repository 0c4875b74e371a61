"""Composite kernels (sums of scaled products of RBF / Matern / RQ / periodic factors: the reference's CO2 model,
experiments/co2_bayesian_sgpr_hmc.py:74-83, :107-152) through the C ABI against oracle/composite.py (torch CPU float64, autograd).

Tolerance: 1e-8 relative (BASELINE.json north_star), max-norm per gradient block.
"""
import os

import pytest
import torch

from helpers import make_problem, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-8
DEV = "cuda:0"
PROG_ALL = (("periodic", "rbf"), ("rq",), ("matern32", "rbf"), ("matern52",))


def _kth(prog, d, seed):
    from oracle import composite as C
    g = torch.Generator().manual_seed(seed)
    P = C.nparams(prog, d)
    kth = 0.7 + 0.8 * torch.rand(P, dtype=torch.float64, generator=g)       # amplitudes, lengthscales, alpha, periods all in [0.7, 1.5]
    return kth


def _engine(prog, **kw):
    import ggp_b200
    assert os.path.exists(ggp_b200.LIB_PATH), "CUDA extension missing"
    return ggp_b200.Engine.get(torch.device(DEV), prog, **kw)


def test_composite_kernel_matrix_matches_the_oracle():
    from oracle import composite as C
    X, y, Z, _ = make_problem(150, 24, 2, seed=1)
    kth = _kth(PROG_ALL, 2, 3)
    eng = _engine(PROG_ALL)
    th = torch.cat([kth, torch.tensor([0.1], dtype=torch.float64)])
    K = eng.kernel_matrix(X, Z, th)
    assert relerr(K, C.composite_kernel(PROG_ALL, kth, X, Z)) < 1e-13
    # a one-term, one-factor program is the single kernel of the existing paths
    import ggp_b200
    for name in ("rbf", "matern32", "matern52"):
        e1 = _engine(((name,),))
        e0 = ggp_b200.Engine.get(torch.device(DEV), name)
        ell, sf2 = torch.tensor([0.9, 1.4], dtype=torch.float64), torch.tensor(1.3, dtype=torch.float64)
        K1 = e1.kernel_matrix(X, Z, torch.cat([sf2.reshape(1), ell, torch.tensor([0.1], dtype=torch.float64)]))
        K0 = e0.kernel_matrix(X, Z, torch.cat([ell, sf2.reshape(1), torch.tensor([0.1], dtype=torch.float64)]))
        assert relerr(K1, K0) < 1e-14, name


@pytest.mark.parametrize("chunk_rows", [0, 256])
def test_composite_bound_and_gradient_match_autograd(chunk_rows):
    """Bound, d/d(kernel parameters), d/ds2, d/dZ for a program with every factor kind, two theta rows in one batched call; with
    chunk_rows=256 the 700 rows stream in three chunks (accumulated row sums, rebuilt tiles)."""
    from oracle import composite as C
    X, y, Z, _ = make_problem(700, 40, 2, seed=5)
    P = C.nparams(PROG_ALL, 2)
    thetas = torch.stack([torch.cat([_kth(PROG_ALL, 2, s), torch.tensor([0.05 + 0.1 * s], dtype=torch.float64)]) for s in (1, 2)])
    eng = _engine(PROG_ALL, chunk_rows=chunk_rows)
    out = eng.sgpr_eval(X.to(DEV), y.to(DEV), Z.to(DEV), thetas.to(DEV), jitter_policy=1e-6)
    assert out["grad"].shape == (2, P + 1 + Z.numel())
    for b in range(2):
        F, g = C.sgpr_bound_and_grads_composite(X, y, Z, PROG_ALL, thetas[b, :P], thetas[b, P], jitter_policy=1e-6, normalize="none")
        assert relerr(out["bound"][b], F) < TOL
        assert relerr(out["grad"][b, :P], g["k"]) < TOL
        assert relerr(out["grad"][b, P], g["s2"]) < TOL
        assert relerr(out["grad"][b, P + 1:].reshape(Z.shape), g["Z"]) < TOL


def test_composite_autograd_function_and_predictive():
    from oracle import composite as C
    import ggp_b200.functions as F
    prog = (("periodic", "rbf"), ("rq",))
    X, y, Z, _ = make_problem(300, 20, 1, seed=8)
    kth = _kth(prog, 1, 4)
    s2 = torch.tensor(0.15, dtype=torch.float64)
    Zp = Z.clone().to(DEV).requires_grad_(True)
    kp = kth.clone().to(DEV).requires_grad_(True)
    sp = s2.clone().to(DEV).requires_grad_(True)
    cfg = dict(kernel=prog, jitter_policy=1e-6)
    loss = -F.sgpr_bound_composite(X.to(DEV), y.to(DEV), Zp, kp, sp, cfg)
    loss.backward()
    Fo, g = C.sgpr_bound_and_grads_composite(X, y, Z, prog, kth, s2, jitter_policy=1e-6, normalize="n")
    assert relerr(-loss, Fo) < TOL
    assert relerr(-kp.grad, g["k"]) < TOL and relerr(-sp.grad, g["s2"]) < TOL and relerr(-Zp.grad, g["Z"]) < TOL
    # eval-mode predictive with the diagonal correction on the training rows (models/sgpr.py:150-160)
    eng = _engine(prog)
    th = torch.cat([kth, s2.reshape(1)])
    Xs = torch.linspace(-2.0, 2.0, 37, dtype=torch.float64).unsqueeze(1)
    for tdc in (True, False):
        eng.sgpr_predict_state(X, y, Z, th, jitter_policy=1e-6, train_diag_correction=tdc)
        mean, var, cov = eng.sgpr_predict(Xs, Z, th, full_cov=True)
        mo, co = C.sgpr_predict_composite(Xs, X, y, Z, prog, kth, s2, jitter_policy=1e-6, train_diag_correction=tdc)
        assert relerr(mean[0], mo) < TOL and relerr(cov[0], co) < TOL and relerr(var[0], torch.diagonal(co)) < TOL


def test_co2_model_logp_dlogp_matches_the_oracle_and_nuts_runs_on_it():
    """The pymc3 CO2 target (11 unconstrained parameters) at BASELINE configs[1]'s shape, 3 chains in one batched call."""
    from oracle import composite as C
    import ggp_b200.functions as F
    import ggp_b200.synthetic as syn
    from ggp_b200.hmc import nuts_sample
    c = syn.config2_co2_shaped(545, 100)
    X, y, Z = (torch.tensor(c[k]) for k in ("X", "y", "Z"))
    g = torch.Generator().manual_seed(2)
    xs = 0.3 * torch.randn(3, 11, dtype=torch.float64, generator=g)
    xs[:, 10] -= 1.0
    Xd, yd, Zd = X.to(DEV), y.to(DEV), Z.to(DEV)
    lp, dlp = F.co2_logp_dlogp(xs.to(DEV), Xd, yd, Zd)
    for cidx in range(3):
        lo, go = C.co2_logp_dlogp(xs[cidx], X, y, Z)
        assert relerr(lp[cidx], lo) < TOL
        assert relerr(dlp[cidx], go) < TOL
    gen = torch.Generator(device=DEV).manual_seed(3)
    res = nuts_sample(lambda xx: F.co2_logp_dlogp(xx, Xd, yd, Zd), xs.to(DEV), 5, tune=10, max_treedepth=4, generator=gen)
    assert torch.isfinite(res["logp"]).all() and res["samples"].shape == (5, 3, 11)
    assert float(res["logp"].mean()) > float(lp.mean())          # the chains climb from the random start


def test_composite_is_rejected_where_it_is_not_implemented():
    from ggp_b200._lib import GgpError
    prog = (("rbf",), ("matern32",))
    X, y, Z, _ = make_problem(128, 16, 2, seed=2)
    eng = _engine(prog)
    th = torch.cat([_kth(prog, 2, 1), torch.tensor([0.1], dtype=torch.float64)])
    with pytest.raises(ValueError):
        eng.sgpr_eval(X, y, Z, th[:-1])                          # wrong row length
    qm = torch.zeros(16, dtype=torch.float64)
    qL = torch.eye(16, dtype=torch.float64)
    with pytest.raises((GgpError, AssertionError, ValueError)):
        eng.svgp_eval(X, y, Z, qm, qL, torch.ones(4, dtype=torch.float64), num_data=128)
