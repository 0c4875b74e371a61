"""GPU parity: whitened SVGP minibatch ELBO (+ all gradients) and the SGPMC log-density against the CPU oracle.
Tolerance 1e-8 relative (float64 mode) on values; gradients 1e-8 at moderate conditioning (see tests/test_gpu_sgpr.py)."""
import os

import numpy as np
import pytest
import torch

from helpers import make_problem, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-8
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def eng():
    import ggp_b200
    return ggp_b200.Engine.get(torch.device("cuda:0"))


def _qu(M, seed):
    rs = np.random.RandomState(seed)
    m = torch.tensor(0.3 * rs.randn(M))
    Ls = torch.tensor(np.eye(M) + 0.1 * rs.randn(M, M))   # raw (unmasked): the kernel must apply the lower mask itself
    return m, Ls


def _oracle(xb, yb, Z, m, Ls, th, N, lik, base, kind="rbf"):
    from oracle import svgp as osv
    D = xb.shape[1]
    old = osv.VAR_CHOL_JITTER_F64
    osv.VAR_CHOL_JITTER_F64 = base
    try:
        ps = [t.clone().requires_grad_(True) for t in (th[:D], th[D], th[D + 1], Z, m, Ls)]
        e = osv.svgp_elbo(xb, yb, ps[3], ps[4], ps[5], ps[0], ps[1], ps[2], N, likelihood=lik, jitter_policy=0.0, kind=kind)
        g = torch.autograd.grad(e, ps, allow_unused=True)
    finally:
        osv.VAR_CHOL_JITTER_F64 = old
    return e.detach(), g


@pytest.mark.parametrize("N,M,D,B,lik", [(400, 30, 3, 64, "gaussian"), (2000, 100, 4, 256, "gaussian"), (9568, 500, 4, 1024, "gaussian"),
                                         (400, 30, 3, 64, "bernoulli"), (3000, 129, 16, 300, "bernoulli")])
def test_svgp_elbo_and_gradients(eng, N, M, D, B, lik):
    X, y, Z, th = make_problem(N, M, D, seed=N + M)
    m, Ls = _qu(M, M)
    xb = X[:B]
    yb = y[:B] if lik == "gaussian" else (y[:B] > 0).double()
    base = 1e-4  # total Kzz jitter (moderate conditioning; the gpytorch default 1e-6 case is test_svgp_default_jitter_and_golden)
    out = eng.svgp_eval(xb, yb, Z, m, Ls, th, num_data=N, likelihood=lik, jitter_policy=0.0, base_jitter=base)
    eo, go = _oracle(xb, yb, Z, m, Ls, th, N, lik, base)
    g = out["grad"][0].cpu()
    o = D + 2
    assert out["info"].tolist() == [0]
    assert relerr(out["value"], eo) < 1e-8
    assert relerr(g[:D], go[0]) < 1e-8
    assert relerr(g[D], go[1]) < 1e-8
    if lik == "gaussian":
        assert relerr(g[D + 1], go[2]) < 1e-8
    assert relerr(g[o:o + M * D].view(M, D), go[3]) < 1e-8
    assert relerr(g[o + M * D:o + M * D + M], go[4]) < 1e-8
    assert relerr(g[o + M * D + M:].view(M, M), torch.tril(go[5])) < 1e-8


def test_svgp_default_jitter_and_golden(eng):
    g = np.load(os.path.join(GOLD, "svgp_small.npz"))
    T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    xb, yb, Z, th, m, Ls = (T(g[k]) for k in ("xb", "yb", "Z", "theta", "m", "Ls"))
    out = eng.svgp_eval(xb, yb, Z, m, Ls, th, num_data=float(g["num_data"]), need_grad=False)   # gpytorch defaults
    assert relerr(out["value"], float(g["elbo_unwhitened"])) < 1e-8
    outb = eng.svgp_eval(xb, T(g["yb01"]), Z, m, Ls, th, num_data=float(g["num_data"]), likelihood="bernoulli", need_grad=False)
    assert relerr(outb["value"], float(g["elbo_bernoulli_gh20"])) < 1e-8


def test_batched_theta_draws_bayesian_svgp(eng):
    """models/bayesian_svgp.py:160-167: 5 theta draws per minibatch share (Z, m, L_s); one batched launch sequence."""
    N, M, D, B = 1500, 64, 3, 200
    X, y, Z, th = make_problem(N, M, D, seed=9)
    m, Ls = _qu(M, 3)
    gen = torch.Generator().manual_seed(1)
    thetas = th.unsqueeze(0) * (0.7 + 0.6 * torch.rand(5, D + 2, dtype=torch.float64, generator=gen))
    outb = eng.svgp_eval(X[:B], y[:B], Z, m, Ls, thetas, num_data=N, jitter_policy=0.0, base_jitter=1e-5)
    for b in range(5):
        o = eng.svgp_eval(X[:B], y[:B], Z, m, Ls, thetas[b], num_data=N, jitter_policy=0.0, base_jitter=1e-5)
        assert torch.equal(o["value"][0], outb["value"][b]) and torch.equal(o["grad"][0], outb["grad"][b])


def test_autograd_function_svgp_step(eng):
    import ggp_b200.functions as F
    N, M, D, B = 800, 40, 2, 128
    X, y, Z, th = make_problem(N, M, D, seed=31)
    m, Ls = _qu(M, 5)
    dev = eng.device
    ps = [t.to(dev).clone().requires_grad_(True) for t in (Z, m, Ls, th[:D], th[D], th[D + 1])]
    loss = -F.svgp_elbo(X[:B].to(dev), y[:B].to(dev), ps[0], ps[1], ps[2], ps[3], ps[4], ps[5], N, dict(jitter_policy=0.0))
    loss.backward()
    eo, go = _oracle(X[:B], y[:B], Z, m, Ls, th, N, "gaussian", 1e-6)
    assert relerr(loss, -eo) < 1e-8
    assert relerr(ps[1].grad, -go[4]) < TOL and relerr(ps[2].grad, -torch.tril(go[5])) < TOL
    assert relerr(ps[3].grad, -go[0]) < TOL and relerr(ps[0].grad, -go[3]) < TOL


@pytest.mark.parametrize("lik,N,M,D", [("gaussian", 700, 40, 3), ("bernoulli", 6000, 70, 16)])
def test_sgpmc_log_density(eng, lik, N, M, D):
    """models/sgp_hmc.py:63: SGPMC.log_posterior_density (whitened v, softplus-raw hyper-parameters, Gamma(2,1) priors)."""
    import ggp_b200.functions as F
    from oracle import sgpmc
    X, y, Z, th = make_problem(N, M, D, seed=N)
    yy = y if lik == "gaussian" else (y > 0).double()
    gen = torch.Generator().manual_seed(2)
    v = 0.5 * torch.randn(2, M, dtype=torch.float64, generator=gen)
    raw = torch.randn(2, D + 2, dtype=torch.float64, generator=gen) * 0.3 + 0.5
    lp, gv, gr = F.sgpmc_logp_dlogp(v, raw, X.to(eng.device), yy.to(eng.device), Z.to(eng.device), likelihood=lik, jitter=1e-4, engine=eng)
    for c in range(2):
        lo, gvo, gro = sgpmc.sgpmc_logp_dlogp(v[c], raw[c], X, yy, Z, likelihood=lik, jitter=1e-4)
        assert relerr(lp[c], lo) < 1e-8
        assert relerr(gv[c], gvo) < 1e-8
        n = D + 2 if lik == "gaussian" else D + 1
        assert relerr(gr[c][:n], gro[:n]) < 1e-8


def test_sgpmc_chain_batch_at_config5_shape(eng):
    """BASELINE configs[4]: Bernoulli-probit classification, N = 2e5, D = 16, M = 512, HMC chains.  One batched launch sequence
    evaluates all chains of a rank (each with its own theta AND its own whitened v); checked against the chunked oracle for two
    of the chains, and chain b of the batch must equal the same chain evaluated alone (bit for bit)."""
    import ggp_b200.functions as F
    import ggp_b200.synthetic as syn
    from oracle import sgpmc
    c = syn.config5_classification()
    X, y, Z = (torch.tensor(c[k]) for k in ("X", "y", "Z"))
    D, M, C = 16, 512, 3
    gen = torch.Generator().manual_seed(5)
    v = 0.3 * torch.randn(C, M, dtype=torch.float64, generator=gen)
    raw = torch.randn(C, D + 2, dtype=torch.float64, generator=gen) * 0.2 + 1.5
    Xd, yd, Zd = X.to(eng.device), y.to(eng.device), Z.to(eng.device)
    lp, gv, gr = F.sgpmc_logp_dlogp(v, raw, Xd, yd, Zd, likelihood="bernoulli", engine=eng)
    lp1, gv1, gr1 = F.sgpmc_logp_dlogp(v[1:2], raw[1:2], Xd, yd, Zd, likelihood="bernoulli", engine=eng)
    assert torch.equal(lp[1], lp1[0]) and torch.equal(gv[1], gv1[0]) and torch.equal(gr[1], gr1[0])
    torch.set_num_threads(__import__("os").cpu_count() or 1)
    for ch in (0, 2):
        lo, gvo, gro = sgpmc.sgpmc_logp_dlogp_chunked(v[ch], raw[ch], X, y, Z, likelihood="bernoulli", chunk=8192)
        assert relerr(lp[ch], lo) < TOL and relerr(gv[ch], gvo) < TOL and relerr(gr[ch][:D + 1], gro[:D + 1]) < TOL, ch


@pytest.mark.parametrize("kind", ["matern32", "matern52", ("rq", 1.5)])
@pytest.mark.parametrize("lik", ["gaussian", "bernoulli"])
def test_svgp_gradients_for_matern_and_rq_kernels(kind, lik):
    """The SVGP / SGPMC backward for the non-RBF tiles: moments weighted by dk/d(d2), the k-weighted total of dF/dsf2 accumulated
    beside them (two chunks of rows here: 4096 + 904)."""
    import ggp_b200
    kw = dict(kernel=kind) if isinstance(kind, str) else dict(kernel=kind[0], kernel_param=kind[1])
    e = ggp_b200.Engine.get(torch.device("cuda:0"), **kw)
    N, M, D, B = 6000, 70, 3, 5000
    X, y, Z, th = make_problem(N, M, D, seed=17)
    m, Ls = _qu(M, M)
    xb = X[:B]
    yb = y[:B] if lik == "gaussian" else (y[:B] > 0).double()
    out = e.svgp_eval(xb, yb, Z, m, Ls, th, num_data=N, likelihood=lik, jitter_policy=0.0, base_jitter=1e-4)
    eo, go = _oracle(xb, yb, Z, m, Ls, th, N, lik, 1e-4, kind=kind)
    g = out["grad"][0].cpu()
    o = D + 2
    assert relerr(out["value"], eo) < TOL
    assert relerr(g[:D], go[0]) < TOL and relerr(g[D], go[1]) < TOL
    if lik == "gaussian":
        assert relerr(g[D + 1], go[2]) < TOL
    assert relerr(g[o:o + M * D].view(M, D), go[3]) < TOL
    assert relerr(g[o + M * D:o + M * D + M], go[4]) < TOL
    assert relerr(g[o + M * D + M:].view(M, M), torch.tril(go[5])) < TOL
