"""GPU parity tests of the sliced-integer tcgen05 path (GGP_PREC_FP64_I8, csrc/gemm_i8.cuh): the same 1e-8 budget as the DMMA path."""
import pytest
import torch

from helpers import make_problem, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-8


@pytest.mark.parametrize("mm,nn,kk", [(128, 64, 64), (200, 100, 70), (1024, 2048, 1024), (384, 384, 4096), (256, 128, 16384), (130, 65, 1000)])
def test_gemm_i8_matches_float64_matmul(mm, nn, kk):
    """Row-scaled 7 balanced radix-256 digits, exact int32 accumulation per significance level: error relative to sum |a||b| at the FP64 level,
    also for rows of very different magnitude (L^{-1} / P like) and ragged shapes (TMA zero fill)."""
    import ggp_b200
    eng = ggp_b200.Engine.get(torch.device("cuda:0"))
    g = torch.Generator().manual_seed(mm + nn)
    A = torch.randn(mm, kk, dtype=torch.float64, generator=g) * torch.exp2(20 * torch.rand(mm, 1, dtype=torch.float64, generator=g) - 10)
    A = A * torch.exp2(-8 * torch.rand(mm, kk, dtype=torch.float64, generator=g))
    B = torch.exp(-6 * torch.rand(nn, kk, dtype=torch.float64, generator=g))
    C = eng.gemm_nt_i8(A, B).cpu()
    ref = A @ B.T
    scale = A.abs() @ B.abs().T
    assert float(((C - ref).abs() / scale).max()) < 2e-15


@pytest.mark.parametrize("N,M,D,jit", [(3000, 260, 4, 1e-4), (2500, 129, 8, 1e-4), (40000, 300, 8, 1e-4), (20000, 1024, 8, 1e-4),
                                       (5000, 200, 16, 1e-4), (1777, 100, 1, 1e-4), (16385, 128, 3, 1e-4)])
def test_i8_bound_and_gradient_parity(N, M, D, jit):
    import ggp_b200
    from oracle import sgpr as osgpr
    eng = ggp_b200.Engine.get(torch.device("cuda:0"), precision="fp64_i8")
    X, y, Z, th = make_problem(N, M, D, seed=N)
    out = eng.sgpr_eval(X, y, Z, th, jitter_policy=jit)
    Fo, go, _ = osgpr.sgpr_bound_and_grads_chunked(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=jit, normalize="none")
    g = out["grad"][0].cpu()
    assert out["info"].tolist() == [0] and out["info_b"].tolist() == [0]
    assert relerr(out["bound"], Fo) < TOL
    assert relerr(g[:D], go["ell"]) < TOL
    assert relerr(g[D], go["sf2"]) < TOL
    assert relerr(g[D + 1], go["s2"]) < TOL
    assert relerr(g[D + 2:].view(M, D), go["Z"]) < TOL


def test_i8_agrees_with_dmma_at_the_conditioning_floor():
    """cond(Kzz) ~ 1e8: both FP64-class paths sit on the same conditioning floor."""
    import ggp_b200
    X, y, Z, th = make_problem(3000, 260, 4, seed=3000)
    a = ggp_b200.Engine.get(torch.device("cuda:0")).sgpr_eval(X, y, Z, th, jitter_policy=1e-6)
    b = ggp_b200.Engine.get(torch.device("cuda:0"), precision="fp64_i8").sgpr_eval(X, y, Z, th, jitter_policy=1e-6)
    assert relerr(b["bound"], a["bound"]) < 1e-10
    assert relerr(b["grad"], a["grad"]) < TOL
    from oracle import hp
    Ft, gt = hp.bound_grad(X.numpy(), y.numpy(), Z.numpy(), th.numpy(), 1e-6, "ld")
    gtv = torch.cat([torch.tensor(gt["ell"]), torch.tensor([gt["sf2"], gt["s2"]]), torch.tensor(gt["Z"]).reshape(-1)])
    for o in (a, b):
        assert relerr(o["bound"], Ft) < TOL and relerr(o["grad"][0], gtv) < TOL


def test_i8_matern_gradient_parity():
    import ggp_b200
    from oracle import sgpr as osgpr
    N, M, D = 2500, 129, 5
    eng = ggp_b200.Engine.get(torch.device("cuda:0"), kernel="matern32", precision="fp64_i8")
    X, y, Z, th = make_problem(N, M, D, seed=N + 1)
    out = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-4)
    Fo, go = osgpr.sgpr_bound_and_grads_autograd(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=1e-4, normalize="none", kind="matern32")
    g = out["grad"][0].cpu()
    assert relerr(out["bound"], Fo) < TOL and relerr(g[:D], go["ell"]) < TOL and relerr(g[D + 2:].view(M, D), go["Z"]) < TOL


def test_i8_row_shard_additivity_and_determinism():
    """Row shards add up (what the NCCL all-reduce does across ranks) and repeated evaluations are bit-identical."""
    import ggp_b200
    eng = ggp_b200.Engine.get(torch.device("cuda:0"), precision="fp64_i8")
    X, y, Z, th = make_problem(30000, 256, 6, seed=12)
    full = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-4)
    again = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-4)
    assert torch.equal(full["bound"], again["bound"]) and torch.equal(full["grad"], again["grad"])
    p1 = eng.sgpr_eval(X[:13000], y[:13000], Z, th, jitter_policy=1e-4)["partial"].clone()
    p2 = eng.sgpr_eval(X[13000:], y[13000:], Z, th, jitter_policy=1e-4)["partial"].clone()
    assert relerr(p1 + p2, full["partial"]) < 1e-13


def _oracle_check(out, X, y, Z, th, jit, M, D):
    from oracle import sgpr as osgpr
    Fo, go, _ = osgpr.sgpr_bound_and_grads_chunked(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=jit, normalize="none")
    g = out["grad"][0].cpu()
    assert out["info"].tolist() == [0] and out["info_b"].tolist() == [0]
    assert relerr(out["bound"], Fo) < TOL
    assert relerr(g[:D], go["ell"]) < TOL and relerr(g[D], go["sf2"]) < TOL and relerr(g[D + 1], go["s2"]) < TOL
    assert relerr(g[D + 2:].view(M, D), go["Z"]) < TOL


def test_i8_execution_plans_agree():
    """The same evaluation through the execution plans of the sliced-integer path:
    (a) tile cache + prefetch on the side stream + one-launch triangular multiply / backward pass (the headline plan; needs
        n >= Engine.prefetch_min_rows), (b) no tile cache: strictly streaming, tiles rebuilt in pass 2, one launch per chunk,
    (c) small ragged chunks (18 chunks, fewer slabs per chunk than CTAs), and the FP64 DMMA path.
    The GPU plans must agree with each other to 1e-9 (two independent implementations: int8-sliced and FP64 DMMA) and with both
    oracle evaluations (chunked and unchunked closed form) to 1e-8."""
    import ggp_b200
    from oracle import sgpr as osgpr
    N, M, D, jit = 70001, 256, 4, 1e-4
    X, y, Z, th = make_problem(N, M, D, seed=77)
    dev = torch.device("cuda:0")
    a = ggp_b200.Engine.get(dev, precision="fp64_i8").sgpr_eval(X, y, Z, th, jitter_policy=jit)
    b = ggp_b200.Engine.get(dev, precision="fp64_i8", tile_cache_mib=0).sgpr_eval(X, y, Z, th, jitter_policy=jit)
    c = ggp_b200.Engine.get(dev, precision="fp64_i8", chunk_rows=4096).sgpr_eval(X, y, Z, th, jitter_policy=jit)
    f = ggp_b200.Engine.get(dev, precision="fp64").sgpr_eval(X, y, Z, th, jitter_policy=jit)
    for o in (b, c, f):
        assert relerr(o["bound"], a["bound"]) < 1e-12
        assert relerr(o["grad"], a["grad"]) < 1e-9
    Fo, go, _ = osgpr.sgpr_bound_and_grads_chunked(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=jit, normalize="none")
    _, gc = osgpr.sgpr_grads_closed_form(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=jit, normalize="none")
    g = a["grad"][0].cpu()
    assert relerr(a["bound"], Fo) < TOL
    for key, mine in (("ell", g[:D]), ("sf2", g[D]), ("s2", g[D + 1]), ("Z", g[D + 2:].view(M, D))):
        assert relerr(mine, go[key]) < TOL and relerr(mine, gc[key]) < TOL, key


def test_i8_host_rows_match_device_rows():
    """Host (pinned and pageable) X, y go up on the side stream next to the factorisation: same bits as device-resident rows."""
    import ggp_b200
    dev = torch.device("cuda:0")
    eng = ggp_b200.Engine.get(dev, precision="fp64_i8")
    X, y, Z, th = make_problem(66000, 128, 3, seed=5)
    ref = eng.sgpr_eval(X.to(dev), y.to(dev), Z.to(dev), th.to(dev), jitter_policy=1e-4)
    for Xh, yh in ((X, y), (X.pin_memory(), y.pin_memory())):
        out = eng.sgpr_eval(Xh, yh, Z, th, jitter_policy=1e-4)
        assert torch.equal(out["bound"], ref["bound"]) and torch.equal(out["grad"], ref["grad"])


def test_cluster_resident_factorisation_plans_agree(monkeypatch):
    """The Kzz factorisation next to the tile build is one thread-block-cluster launch (k_chol_cluster); with a long build beside it the
    explicit inverse rides in the same launch.  Both plans against the oracle at 1e-8 and against each other (same arithmetic per block,
    different summation order inside the inverse: rounding level), also with a failing pivot (LAPACK-style info through the cluster launch)."""
    import ggp_b200
    from oracle import sgpr as osgpr
    dev = torch.device("cuda:0")
    eng = ggp_b200.Engine.get(dev, precision="fp64_i8")
    N, M, D = 70000, 300, 4          # padded M = 512: 8 diagonal blocks, 3 inverse levels; N >= prefetch_min_rows: the side-stream build
    X, y, Z, th = make_problem(N, M, D, seed=11)
    Xd, yd, Zd, thd = X.to(dev), y.to(dev), Z.to(dev), th.to(dev)
    Fo, go = osgpr.sgpr_grads_closed_form(X, y, Z, th[:D], th[D], th[D + 1], jitter_policy=1e-4, normalize="none")
    outs = {}
    for plan, env in (("inverse in the cluster launch", {"GGP_CHOL_CLUSTER_INV": "1"}), ("inverse as launches", {"GGP_CHOL_CLUSTER_NO_INV": "1"})):
        for k in ("GGP_CHOL_CLUSTER_INV", "GGP_CHOL_CLUSTER_NO_INV"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        out = eng.sgpr_eval(Xd, yd, Zd, thd, jitter_policy=1e-4)
        g = out["grad"][0].cpu()
        assert relerr(out["bound"], Fo) < TOL, plan
        for key, mine in (("ell", g[:D]), ("sf2", g[D]), ("s2", g[D + 1]), ("Z", g[D + 2:].view(M, D))):
            assert relerr(mine, go[key]) < TOL, (plan, key)
        outs[plan] = out
    a, b = outs.values()
    assert relerr(a["bound"], b["bound"]) < 1e-13 and relerr(a["grad"], b["grad"]) < 1e-10
    # a Kzz that is not positive definite at jitter 0 (duplicated inducing rows): the ladder engages through the cluster launch's info
    monkeypatch.setenv("GGP_CHOL_CLUSTER_INV", "1")
    monkeypatch.delenv("GGP_CHOL_CLUSTER_NO_INV", raising=False)
    Zdup = Zd.clone()
    Zdup[200] = Zdup[3]
    out = eng.sgpr_eval(Xd, yd, Zdup, thd, jitter_policy="gpytorch")
    assert float(out["jitter"][0]) > 0.0 and bool(torch.isfinite(out["bound"]).all())


@pytest.mark.parametrize("D", [11, 12])
def test_i8_moment_block_boundary(D):
    """2 D + 1 = 23 moments still fit the register-resident accumulation (three DMMA blocks); D = 12 takes the per-tile path."""
    import ggp_b200
    N, M, jit = 6000, 200, 1e-4
    X, y, Z, th = make_problem(N, M, D, seed=D)
    out = ggp_b200.Engine.get(torch.device("cuda:0"), precision="fp64_i8").sgpr_eval(X, y, Z, th, jitter_policy=jit)
    _oracle_check(out, X, y, Z, th, jit, M, D)


def test_i8_batch_of_theta_rows_runs_draw_by_draw():
    """batch > 1 on a large streamed problem (the |trace| hyper-parameter draws of models/bayesian_sgpr_hmc.py:121-134): every row
    equals its single evaluation bit for bit, on the tcgen05 path."""
    import ggp_b200
    eng = ggp_b200.Engine.get(torch.device("cuda:0"), precision="fp64_i8")
    X, y, Z, th = make_problem(140000, 512, 4, seed=9)
    assert X.shape[0] * 512 >= eng.i8_batch_min_elems
    g = torch.Generator().manual_seed(1)
    thetas = th.unsqueeze(0) * (0.8 + 0.4 * torch.rand(3, 6, dtype=torch.float64, generator=g))
    dev = eng.device
    Xd, yd = X.to(dev), y.to(dev)
    outb = eng.sgpr_eval(Xd, yd, Z, thetas, jitter_policy=1e-4)
    assert outb["path"] == "fp64_i8" and outb["bound"].shape == (3,) and outb["grad"].shape == (3, 6 + 512 * 4)
    for b in range(3):
        o = eng.sgpr_eval(Xd, yd, Z, thetas[b], jitter_policy=1e-4)
        assert torch.equal(o["bound"][0], outb["bound"][b]) and torch.equal(o["grad"][0], outb["grad"][b])
    f = ggp_b200.Engine.get(dev, precision="fp64").sgpr_eval(Xd, yd, Z, thetas, jitter_policy=1e-4)
    assert relerr(outb["bound"], f["bound"]) < 1e-12 and relerr(outb["grad"], f["grad"]) < 1e-9
