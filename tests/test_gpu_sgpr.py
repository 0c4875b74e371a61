"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors.

Tolerances (BASELINE.json north_star): 1e-8 relative in float64 on the bound, its gradients and the predictive.
Gradient parity is asserted at moderate conditioning of Kzz; at cond(Kzz) ~ 1e8 the oracle's own two float64 gradient
evaluations disagree above 1e-9 (tests/test_oracle.py::test_gradient_conditioning_floor_is_inherent), so the
ill-conditioned case is asserted -- at the same 1e-8 -- against the long-double evaluation of oracle/hp.
"""
import os

import numpy as np
import pytest
import torch

from helpers import make_problem, relerr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-8


@pytest.fixture(scope="module")
def eng():
    import ggp_b200
    assert os.path.exists(ggp_b200.LIB_PATH), "CUDA extension missing"
    return ggp_b200.Engine.get(torch.device("cuda:0"))


def test_dmma_gemm_matches_torch(eng):
    dev = eng.device
    g = torch.Generator(device="cpu").manual_seed(0)
    for (mm, nn, kk) in [(128, 128, 64), (200, 70, 34), (512, 384, 1000), (1, 1, 2), (130, 257, 18)]:
        A = torch.randn(mm, kk, dtype=torch.float64, generator=g).to(dev)
        B = torch.randn(nn, kk, dtype=torch.float64, generator=g).to(dev)
        C = eng.gemm_nt(A, B)
        ref = A @ B.T
        assert relerr(C, ref) < 1e-13, (mm, nn, kk)
        C2 = eng.gemm_nt(A, B, C=ref.clone(), alpha=-0.5, beta=2.0)
        assert relerr(C2, 1.5 * ref) < 1e-13


def test_gemm_tile_kernels_and_triangular_k_ranges(eng):
    """Both families of the DMMA product: the 64 x 64-tile kernel of the m x m section (k_mm64: small outputs, skinny outputs) and the
    persistent 128 x 128 kernels (more than 96 work items, or operands that are not 16-byte aligned), with every triangular k-range
    mode the library uses (the clipped ranges only skip exact zeros, so the result is the plain product of the triangular operands)."""
    dev = eng.device
    g = torch.Generator(device="cpu").manual_seed(5)
    KM_A_LOWER, KM_A_UPPER, KM_B_LOWER, KM_B_UPPER = 1, 2, 4, 8
    for n in (300, 1024, 1600):            # 1600: 13 x 13 = 169 work items of 128 x 128 -> the persistent kernel
        F = torch.randn(n, n, dtype=torch.float64, generator=g).to(dev)
        Lo, Up = torch.tril(F), torch.triu(F)
        for a, b, km in ((Lo, F, KM_A_LOWER), (Up, F, KM_A_UPPER), (F, Lo, KM_B_LOWER), (F, Up, KM_B_UPPER), (Up, Up, KM_A_UPPER | KM_B_UPPER),
                         (Lo, Lo, KM_A_LOWER | KM_B_LOWER), (F, F, 0)):
            C = torch.full((n, n), 7.0, dtype=torch.float64, device=dev)
            eng.gemm_nt_ex(a, b, C, kmode=km)
            assert relerr(C, a @ b.T) < 1e-13, (n, km)
    # skinny output (always on the 64-wide tiles), exact tile multiples, ragged edges; accumulate into C
    for (mm, nn, kk) in [(200, 70, 34), (500, 17, 4096), (64, 64, 64), (65, 63, 130)]:
        A = torch.randn(mm, kk, dtype=torch.float64, generator=g).to(dev)
        B = torch.randn(nn, kk, dtype=torch.float64, generator=g).to(dev)
        ref = A @ B.T
        assert relerr(eng.gemm_nt(A, B), ref) < 1e-13, (mm, nn, kk)
        assert relerr(eng.gemm_nt(A, B, C=ref.clone(), alpha=1.0, beta=1.0), 2.0 * ref) < 1e-13, (mm, nn, kk)
    # rows that are not 16-byte aligned are refused loudly (the operand loads are 16-byte LDGSTS / TMA)
    import ggp_b200
    with pytest.raises(ggp_b200.GgpError):
        eng.gemm_nt(torch.randn(8, 33, dtype=torch.float64, device=dev), torch.randn(8, 33, dtype=torch.float64, device=dev))


def test_batched_cholesky_and_inverse(eng):
    dev = eng.device
    g = torch.Generator().manual_seed(1)
    for m in [1, 20, 64, 65, 100, 128, 129, 500, 1000]:
        R = torch.randn(3, m, m, dtype=torch.float64, generator=g).to(dev)
        S = R @ R.transpose(1, 2) + m * torch.eye(m, dtype=torch.float64, device=dev)
        L, Linv, info = eng.chol(S)
        Lt = torch.linalg.cholesky(S)
        assert info.tolist() == [0, 0, 0]
        assert relerr(torch.tril(L), Lt) < 1e-13
        assert relerr(Linv, torch.linalg.inv(Lt)) < 1e-12
    # LAPACK-style failure index
    bad = torch.eye(70, dtype=torch.float64, device=dev).repeat(2, 1, 1)
    bad[1, 66, 66] = -1.0
    _, _, info = eng.chol(bad)
    assert info.tolist() == [0, 67]
    # ... also when the failing block is factored by the look-ahead tail of a trailing-update launch (block 3 of 5), first failure wins
    bad = torch.eye(300, dtype=torch.float64, device=dev).repeat(2, 1, 1)
    bad[0, 200, 200] = -2.0
    bad[0, 290, 290] = -1.0
    _, _, info = eng.chol(bad)
    assert info.tolist() == [201, 0]


@pytest.mark.parametrize("kind", ["rbf", "matern32", "matern52"])
def test_kernel_tiles_match_oracle(kind):
    import ggp_b200
    from oracle.kernels import ard_kernel
    e = ggp_b200.Engine.get(torch.device("cuda:0"), kernel=kind)
    for (n1, n2, d) in [(64, 64, 8), (130, 77, 3), (1, 5, 1), (257, 64, 16)]:
        g = torch.Generator().manual_seed(n1)
        X1 = torch.randn(n1, d, dtype=torch.float64, generator=g)
        X2 = torch.randn(n2, d, dtype=torch.float64, generator=g)
        th = torch.cat([0.5 + torch.rand(d, dtype=torch.float64, generator=g), torch.tensor([1.7, 0.1], dtype=torch.float64)])
        K = e.kernel_matrix(X1, X2, th)
        ref = ard_kernel(X1, X2, th[:d], th[d], kind)
        assert relerr(K, ref) < 1e-13, (kind, n1, n2, d)


@pytest.mark.parametrize("N,M,D,jit", [(300, 20, 1, 1e-4), (1000, 100, 3, 1e-4), (3000, 260, 4, 1e-4), (2500, 129, 8, 1e-4),
                                       (40000, 300, 8, 1e-4)])
def test_bound_and_gradient_parity(eng, N, M, D, jit):
    from oracle import sgpr as osgpr
    X, y, Z, th = make_problem(N, M, D, seed=N)
    out = eng.sgpr_eval(X, y, Z, th, jitter_policy=jit)
    if N <= 5000:
        Fo, go = osgpr.sgpr_grads_closed_form(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=jit, normalize="none")
    else:
        Fo, go, _ = osgpr.sgpr_bound_and_grads_chunked(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=jit, normalize="none")
    g = out["grad"][0].cpu()
    assert out["info"].tolist() == [0] and out["info_b"].tolist() == [0]
    assert relerr(out["bound"], Fo) < TOL
    assert relerr(g[:D], go["ell"]) < TOL
    assert relerr(g[D], go["sf2"]) < TOL
    assert relerr(g[D + 1], go["s2"]) < TOL
    assert relerr(g[D + 2:].view(M, D), go["Z"]) < TOL


@pytest.mark.parametrize("kind", ["matern32", "matern52", ("rq", 0.7), ("rq", 3.0)])
@pytest.mark.parametrize("N,M,D", [(700, 40, 2), (2500, 129, 5)])
def test_matern_and_rq_bound_and_gradient_parity(kind, N, M, D):
    """Matern-3/2, -5/2 and rational-quadratic tiles (experiments/co2_bayesian_sgpr_hmc.py:74-83,127-144 uses Matern32 and RatQuad):
    bound and analytic gradient (the backward epilogue weights G by dk/d(d2); dF/d sf2 comes from the m x m section) against oracle
    autograd.  The RQ shape parameter alpha is a constant of the evaluation (cfg.kernel_param)."""
    import ggp_b200
    from oracle import sgpr as osgpr
    kw = dict(kernel=kind) if isinstance(kind, str) else dict(kernel=kind[0], kernel_param=kind[1])
    e = ggp_b200.Engine.get(torch.device("cuda:0"), **kw)
    X, y, Z, th = make_problem(N, M, D, seed=N + 1)
    out = e.sgpr_eval(X, y, Z, th, jitter_policy=1e-4)
    Fo, go = osgpr.sgpr_bound_and_grads_autograd(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=1e-4, normalize="none", kind=kind)
    g = out["grad"][0].cpu()
    assert relerr(out["bound"], Fo) < TOL
    assert relerr(g[:D], go["ell"]) < TOL
    assert relerr(g[D], go["sf2"]) < TOL
    assert relerr(g[D + 1], go["s2"]) < TOL
    assert relerr(g[D + 2:].view(M, D), go["Z"]) < TOL


def test_ill_conditioned_parity_at_the_float64_floor(eng):
    from oracle import sgpr as osgpr
    N, M, D = 3000, 260, 4
    X, y, Z, th = make_problem(N, M, D, seed=N)
    out = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-6)  # cond(Kzz) ~ 1e8
    # reference: the long-double evaluation (oracle/hp), whose own error at this conditioning is ~1e-11; the float64 oracle is
    # within 1e-8 of it too (tests/test_oracle_hp.py::test_conditioning_floor_of_the_two_backward_forms)
    from oracle import hp
    Ft, gt = hp.bound_grad(X.numpy(), y.numpy(), Z.numpy(), th.numpy(), 1e-6, "ld")
    g = out["grad"][0].cpu()
    assert relerr(out["bound"], Ft) < TOL
    assert relerr(g[:D], gt["ell"]) < TOL and relerr(g[D], gt["sf2"]) < TOL and relerr(g[D + 1], gt["s2"]) < TOL
    assert relerr(g[D + 2:].view(M, D), gt["Z"]) < TOL


@pytest.mark.parametrize("name", ["sgpr_small_1d", "sgpr_small_3d", "sgpr_mid_4d"])
def test_golden_dense_definition(eng, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    X, y, Z, th = T(g["X"]), T(g["y"]), T(g["Z"]), T(g["theta"])
    M, D = Z.shape
    out = eng.sgpr_eval(X, y, Z, th, jitter_policy=float(g["jitter"]))
    assert relerr(out["bound"], float(g["F_dense"])) < TOL
    gr = out["grad"][0].cpu()
    assert relerr(gr[:D], g["g_ell"]) < TOL and relerr(gr[D + 1], g["g_s2"]) < TOL
    assert relerr(gr[D + 2:].view(M, D), g["g_Z"]) < TOL
    # state left by the bound evaluation: plain noise s2 on the training rows
    mean, var, cov = eng.sgpr_predict(T(g["Xs"]), Z, th, full_cov=True)
    assert relerr(mean[0], g["pred_mean_plain"]) < TOL
    assert relerr(cov[0], g["pred_cov_plain"]) < TOL
    # gpytorch's eval mode: the diagonal correction also sits on the training rows (Lambda_n = s2 + max(k_nn - q_nn, 0))
    eng.sgpr_predict_state(X, y, Z, th, jitter_policy=float(g["jitter"]))
    mean, var, cov = eng.sgpr_predict(T(g["Xs"]), Z, th, full_cov=True)
    assert relerr(mean[0], g["pred_mean"]) < TOL
    assert relerr(cov[0], g["pred_cov"]) < TOL
    assert relerr(var[0], np.diag(g["pred_cov"])) < TOL


def test_predictive_full_covariance_beyond_one_chunk(eng):
    """More test rows than the handle's chunk (small training sets reserve small chunks): the covariance is tiled."""
    from oracle import sgpr as osgpr
    X, y, Z, th = make_problem(300, 40, 2, seed=21)
    Xs = torch.tensor(np.random.RandomState(5).randn(1100, 2))
    eng.sgpr_predict_state(X, y, Z, th, jitter_policy=1e-6)
    mean, var, cov = eng.sgpr_predict(Xs, Z, th, full_cov=True)
    mo, co = osgpr.sgpr_predict(Xs, X, y, Z, th[:2], th[2], th[3], jitter_policy=1e-6)
    assert relerr(mean[0], mo) < TOL and relerr(cov[0], co) < TOL and relerr(var[0], torch.diagonal(co)) < TOL


def test_jitter_ladder_with_duplicate_inducing_rows(eng):
    """Z drawn WITH replacement (experiments/regression.py:83) makes Kzz exactly singular: the gpytorch ladder must engage
    and pick the same jitter as the oracle's psd_safe_cholesky."""
    from oracle import sgpr as osgpr
    from oracle.linalg import psd_safe_cholesky
    from oracle.kernels import ard_kernel
    X, y, Z, th = make_problem(2000, 120, 3, seed=77, without_replacement=False)
    Z[7] = Z[3]
    out = eng.sgpr_eval(X, y, Z, th, jitter_policy="gpytorch")
    _, jit = psd_safe_cholesky(ard_kernel(Z, Z, th[:3], th[3]), "gpytorch")
    assert jit > 0 and abs(float(out["jitter"][0]) - jit) < 1e-20
    Fo = osgpr.sgpr_bound(X, y, Z, th[:3], th[3], th[4], "gpytorch", "none")
    assert relerr(out["bound"], Fo) < TOL
    import ggp_b200
    with pytest.raises(ggp_b200.NotPSDError):
        eng.sgpr_eval(X, y, Z, th, jitter_policy=0.0)


def test_batched_theta_rows_equal_single_evaluations(eng):
    N, M, D = 1500, 64, 3
    X, y, Z, th = make_problem(N, M, D, seed=5)
    g = torch.Generator().manual_seed(3)
    thetas = th.unsqueeze(0) * (0.7 + 0.6 * torch.rand(5, D + 2, dtype=torch.float64, generator=g))
    outb = eng.sgpr_eval(X, y, Z, thetas, jitter_policy=1e-5)
    for b in range(5):
        o = eng.sgpr_eval(X, y, Z, thetas[b], jitter_policy=1e-5)
        assert torch.equal(o["bound"][0], outb["bound"][b])
        assert torch.equal(o["grad"][0], outb["grad"][b])


def test_ragged_and_tiny_shapes(eng):
    from oracle import sgpr as osgpr
    for (N, M, D) in [(1, 1, 1), (7, 3, 2), (129, 65, 5), (1025, 17, 2)]:
        X, y, Z, th = make_problem(max(N, 4), M, D, seed=N + 100)
        X, y = X[:N], y[:N]
        out = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-4)
        Fo, go = osgpr.sgpr_grads_closed_form(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=1e-4, normalize="none")
        assert relerr(out["bound"], Fo) < TOL, (N, M, D)
        assert relerr(out["grad"][0][D + 2:].cpu().view(M, D), go["Z"]) < TOL, (N, M, D)


def test_shard_additivity_and_determinism_at_scale(eng):
    """Size-independent properties at a BASELINE-scale M: pass-1 partial sums over two row shards add to the unsharded
    partial (what the NCCL all-reduce relies on), repeated evaluations are bit-identical, and the analytic gradient
    matches a central finite difference of the bound along a random direction."""
    import ctypes
    N, M, D = 200_000, 1024, 8
    X, y, Z, th = make_problem(N, M, D, seed=1)
    dev = eng.device
    X, y, Z, th = X.to(dev), y.to(dev), Z.to(dev), th.to(dev)
    full = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-6)
    again = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-6)
    assert torch.equal(full["bound"], again["bound"]) and torch.equal(full["grad"], again["grad"])
    h = N // 2 + 37
    p1 = eng.sgpr_eval(X[:h], y[:h], Z, th, jitter_policy=1e-6, need_grad=False)["partial"]
    p2 = eng.sgpr_eval(X[h:], y[h:], Z, th, jitter_policy=1e-6, need_grad=False)["partial"]
    assert relerr(p1 + p2, full["partial"]) < 1e-12
    # directional finite difference in log-theta
    g = torch.Generator().manual_seed(2)
    dirn = torch.randn(D + 2, dtype=torch.float64, generator=g).to(dev) * 0.1
    eps = 1e-5
    Fp = eng.sgpr_eval(X, y, Z, th * torch.exp(eps * dirn), jitter_policy=1e-6, need_grad=False)["bound"][0]
    Fm = eng.sgpr_eval(X, y, Z, th * torch.exp(-eps * dirn), jitter_policy=1e-6, need_grad=False)["bound"][0]
    fd = (Fp - Fm) / (2 * eps)
    an = (full["grad"][0][:D + 2] * th * dirn).sum()
    assert abs(fd - an) < 1e-5 * abs(an)


def test_tile_cache_equals_streaming_rebuild(eng):
    """cfg.tile_cache_mib: keeping the k(X,Z) tiles of pass 1 in HBM for pass 2 must be bit-identical to rebuilding them
    (strictly streaming mode), for one and for several theta rows, and must not leak across evaluations with a new theta."""
    import ggp_b200
    N, M, D = 40_000, 200, 5
    X, y, Z, th = make_problem(N, M, D, seed=4)
    stream_eng = ggp_b200.Engine.get(eng.device, tile_cache_mib=0)
    cache_eng = ggp_b200.Engine.get(eng.device, tile_cache_mib=4096)
    ths = torch.stack([th, th * 1.3, th * 0.8])
    for t in (th, ths):
        a = stream_eng.sgpr_eval(X, y, Z, t, jitter_policy=1e-6)
        b = cache_eng.sgpr_eval(X, y, Z, t, jitter_policy=1e-6)
        assert torch.equal(a["bound"], b["bound"]) and torch.equal(a["grad"], b["grad"])
    th2 = th * 1.1
    b2 = cache_eng.sgpr_eval(X, y, Z, th2, jitter_policy=1e-6)
    a2 = stream_eng.sgpr_eval(X, y, Z, th2, jitter_policy=1e-6)
    assert torch.equal(a2["grad"], b2["grad"])


def test_autograd_function_drops_into_a_training_step(eng):
    """SGPRBound.apply replaces forward + mll + backward of models/sgpr.py:123-129 (gpytorch convention: / N, softplus raws)."""
    import ggp_b200.functions as F
    from oracle import sgpr as osgpr
    N, M, D = 800, 40, 2
    X, y, Z, th = make_problem(N, M, D, seed=21)
    dev = eng.device
    raw_l = torch.zeros(1, D, dtype=torch.float64, device=dev, requires_grad=True)
    raw_o = torch.zeros((), dtype=torch.float64, device=dev, requires_grad=True)
    raw_n = torch.zeros(1, dtype=torch.float64, device=dev, requires_grad=True)
    Zp = Z.to(dev).clone().requires_grad_(True)
    sp = torch.nn.functional.softplus
    loss = -F.sgpr_bound(X.to(dev), y.to(dev), Zp, sp(raw_l), sp(raw_o), sp(raw_n) + 1e-4, dict(jitter_policy=1e-5))
    loss.backward()
    # oracle: same chain on CPU
    rl = torch.zeros(1, D, dtype=torch.float64, requires_grad=True)
    ro = torch.zeros((), dtype=torch.float64, requires_grad=True)
    rn = torch.zeros(1, dtype=torch.float64, requires_grad=True)
    Zc = Z.clone().requires_grad_(True)
    lo = -osgpr.sgpr_bound(X, y, Zc, sp(rl).reshape(-1), sp(ro), (sp(rn) + 1e-4).reshape(()), 1e-5, "n")
    lo.backward()
    assert relerr(loss, lo) < TOL
    assert relerr(raw_l.grad, rl.grad) < TOL and relerr(raw_o.grad, ro.grad) < TOL and relerr(raw_n.grad, rn.grad) < TOL
    assert relerr(Zp.grad, Zc.grad) < TOL


def test_batched_pymc3_logp_dlogp(eng):
    import ggp_b200.functions as F
    from oracle import priors
    N, M, D = 545, 100, 1
    import ggp_b200.synthetic as syn
    c = syn.config2_co2_shaped(N, M)
    X, y, Z = (torch.tensor(c[k]) for k in ("X", "y", "Z"))
    g = torch.Generator().manual_seed(4)
    xs = torch.randn(4, D + 2, dtype=torch.float64, generator=g) * 0.3 + torch.tensor([0.0, 0.0, -1.0], dtype=torch.float64)
    lp, dlp = F.sgpr_vfe_logp_dlogp(xs.to(eng.device), X.to(eng.device), y.to(eng.device), Z.to(eng.device))
    for cidx in range(4):
        lo, go = priors.sgpr_vfe_logp_dlogp(xs[cidx], X, y, Z)
        assert relerr(lp[cidx], lo) < TOL    # duplicate Z rows + pymc3's 1e-6 stabilise jitter: cond(Kzz) ~ 1e8
        assert relerr(dlp[cidx], go) < TOL
    # the fused transform / prior kernels (k_vfe_theta, k_vfe_logp) against the torch spelling of the same target (_vfe_eval), with
    # and without the priors, and the failure convention: a non-finite row gives logp = -inf and a zero gradient, others untouched
    dev = eng.device
    for with_prior in (True, False):
        l1, d1 = F.sgpr_vfe_logp_dlogp(xs.to(dev), X.to(dev), y.to(dev), Z.to(dev), with_prior=with_prior)
        l2, d2, _, bad = F._vfe_eval(xs.to(dev), X.to(dev), y.to(dev), Z.to(dev), "pymc3", eng, False, with_prior)
        assert not bool(bad.any()) and relerr(l1, l2) < 1e-13 and relerr(d1, d2) < 1e-13
    xb = xs.clone()
    xb[2, 0] = float("nan")
    lb, db = F.sgpr_vfe_logp_dlogp(xb.to(dev), X.to(dev), y.to(dev), Z.to(dev))
    assert lb[2].item() == -float("inf") and float(db[2].abs().max()) == 0.0
    keep = [0, 1, 3]
    assert torch.equal(lb[keep], lp[keep]) and torch.equal(db[keep], dlp[keep])


def test_predict_refuses_a_clobbered_handle_state(eng):
    """The SVGP / SGPMC scratch aliases the m x m state the sparse predictive reads: after such a call (or at another theta)
    sgpr_predict must fail loudly instead of returning numbers from someone else's factorisation."""
    X, y, Z, th = make_problem(400, 24, 2, seed=3)
    Xs = X[:10]
    eng.sgpr_predict_state(X, y, Z, th, jitter_policy=1e-6)
    m0, v0, _ = eng.sgpr_predict(Xs, Z, th)
    with pytest.raises(RuntimeError):
        eng.sgpr_predict(Xs, Z, th * 1.1)
    eng.svgp_eval(X[:64], y[:64], Z, torch.zeros(24, dtype=torch.float64), torch.eye(24, dtype=torch.float64), th, num_data=400)
    with pytest.raises(RuntimeError):
        eng.sgpr_predict(Xs, Z, th)
    eng.sgpr_predict_state(X, y, Z, th, jitter_policy=1e-6)
    m1, v1, _ = eng.sgpr_predict(Xs, Z, th)
    assert torch.equal(m0, m1) and torch.equal(v0, v1)
