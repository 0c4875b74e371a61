"""torch-facing entry points of the hot path (Python level).

  SGPRBound.apply(X, y, Z, lengthscale, outputscale, noise, cfg)  -> scalar bound (differentiable in Z, ell, sf2, s2)
        drop-in for   output = self.forward(train_x); -mll(output, train_y)     models/sgpr.py:123-129
  sgpr_vfe_logp_dlogp(x[C, D+2], X, y, Z)                          -> (logp[C], dlogp[C, D+2])
        drop-in for the pymc3 model logp/dlogp NUTS evaluates per leapfrog        models/bayesian_sgpr_hmc.py:60-78
        (all C hyper-parameter draws / chains go through the kernels in one batched launch sequence)
"""
import math

import torch

from .engine import Engine

DEFAULT_CFG = dict(normalize="n", jitter_policy="gpytorch", precision="auto", kernel="rbf", group=False, chunk_rows=0)


def pick_precision(n, m, d):
    """precision="auto": the sliced-integer tcgen05 plans (FP64-class accuracy, 2-3x the FP64 DMMA rate) for large streamed problems
    they support (65 <= m <= 4096 inducing points, d <= 16), the FP64 DMMA plans for everything else (small problems are
    latency-bound; both meet the same parity tests)."""
    return "fp64_i8" if (n * m >= (1 << 26) and 65 <= m <= 4096 and d <= 16) else "fp64"


def _cfg(cfg):
    c = dict(DEFAULT_CFG)
    if cfg:
        c.update(cfg)
    return c


class SGPRBound(torch.autograd.Function):
    """Collapsed Titsias bound.  Gradients are w.r.t. the CONSTRAINED values (ell, sf2, s2, Z); the softplus / log chain
    to raw parameters stays in torch autograd (gpytorch raw-parameter convention)."""

    @staticmethod
    def forward(ctx, X, y, Z, lengthscale, outputscale, noise, cfg=None):
        c = _cfg(cfg)
        prec = pick_precision(X.shape[0], Z.shape[0], X.shape[1]) if c["precision"] == "auto" else c["precision"]
        eng = Engine.get(X.device, c["kernel"], prec, c["chunk_rows"])
        D = X.shape[1]
        theta = torch.cat([lengthscale.reshape(-1), outputscale.reshape(-1), noise.reshape(-1)]).to(torch.float64)
        need = any(ctx.needs_input_grad[2:6])
        out = eng.sgpr_eval(X, y, Z, theta, jitter_policy=c["jitter_policy"], need_grad=need, group=c["group"])
        n_total = out["n_total"][0]
        scale = (1.0 / n_total) if c["normalize"] == "n" else torch.ones((), dtype=torch.float64, device=X.device)
        ctx.D, ctx.M = D, Z.shape[0]
        ctx.shapes = (Z.shape, lengthscale.shape, outputscale.shape, noise.shape)
        ctx.dtypes = (Z.dtype, lengthscale.dtype, outputscale.dtype, noise.dtype)
        ctx.save_for_backward(out["grad"][0] * scale if need else torch.empty(0, device=X.device))
        ctx.jitter = out["jitter"]
        return (out["bound"][0] * scale).to(X.dtype)

    @staticmethod
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        D, M = ctx.D, ctx.M
        zs, ls, os_, ns = ctx.shapes
        zd, ld, od, nd = ctx.dtypes
        gout = gout.to(torch.float64)
        gZ = (gout * g[D + 2:].view(M, D)).reshape(zs).to(zd) if ctx.needs_input_grad[2] else None
        gl = (gout * g[:D]).reshape(ls).to(ld) if ctx.needs_input_grad[3] else None
        go = (gout * g[D]).reshape(os_).to(od) if ctx.needs_input_grad[4] else None
        gn = (gout * g[D + 1]).reshape(ns).to(nd) if ctx.needs_input_grad[5] else None
        return None, None, gZ, gl, go, gn, None


def sgpr_bound(X, y, Z, lengthscale, outputscale, noise, cfg=None):
    return SGPRBound.apply(X, y, Z, lengthscale, outputscale, noise, cfg)


class SGPRBoundComposite(torch.autograd.Function):
    """Collapsed bound with a composite kernel -- sums of scaled products of RBF / Matern / RQ / periodic factors, the structure of the
    reference's CO2 model (experiments/co2_bayesian_sgpr_hmc.py:74-83: ScaleKernel(Periodic * RBF) + ScaleKernel(RBF) +
    ScaleKernel(RQ) + ScaleKernel(RBF) inside InducingPointKernel).  cfg["kernel"] is the program (tuple of terms, each a tuple of
    factor names); kparams[P] is its parameter row in program order (a_t, then per factor ell[d], then rq: alpha | periodic:
    period[d]; _lib.make_kprog).  Differentiable in Z, kparams and noise (constrained values)."""

    @staticmethod
    def forward(ctx, X, y, Z, kparams, noise, cfg=None):
        c = _cfg(cfg)
        if isinstance(c["kernel"], str):
            raise ValueError("SGPRBoundComposite: cfg['kernel'] must be a composite program, e.g. (('periodic', 'rbf'), ('rq',))")
        eng = Engine.get(X.device, c["kernel"], "fp64", c["chunk_rows"])
        theta = torch.cat([kparams.reshape(-1), noise.reshape(-1)]).to(torch.float64)
        need = any(ctx.needs_input_grad[2:5])
        out = eng.sgpr_eval(X, y, Z, theta, jitter_policy=c["jitter_policy"], need_grad=need, group=c["group"])
        n_total = out["n_total"][0]
        scale = (1.0 / n_total) if c["normalize"] == "n" else torch.ones((), dtype=torch.float64, device=X.device)
        ctx.P = kparams.numel()
        ctx.shapes = (Z.shape, kparams.shape, noise.shape)
        ctx.dtypes = (Z.dtype, kparams.dtype, noise.dtype)
        ctx.save_for_backward(out["grad"][0] * scale if need else torch.empty(0, device=X.device))
        return (out["bound"][0] * scale).to(X.dtype)

    @staticmethod
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        P = ctx.P
        zs, ks, ns = ctx.shapes
        zd, kd, nd = ctx.dtypes
        gout = gout.to(torch.float64)
        gZ = (gout * g[P + 1:]).reshape(zs).to(zd) if ctx.needs_input_grad[2] else None
        gk = (gout * g[:P]).reshape(ks).to(kd) if ctx.needs_input_grad[3] else None
        gn = (gout * g[P]).reshape(ns).to(nd) if ctx.needs_input_grad[4] else None
        return None, None, gZ, gk, gn, None


def sgpr_bound_composite(X, y, Z, kparams, noise, cfg):
    return SGPRBoundComposite.apply(X, y, Z, kparams, noise, cfg)


LOG2 = math.log(2.0)
LOGPI = math.log(math.pi)
LOG2PI = math.log(2.0 * math.pi)

# ---- the pymc3 CO2 model (experiments/co2_bayesian_sgpr_hmc.py:107-152) -------------------------------------------------------------
# cov = n_per^2 Periodic(1, period=1, ls=l_psmooth) ExpQuad(1, l_pdecay) + n_med^2 RatQuad(1, l_med, alpha) + n_trend^2 ExpQuad(1, l_trend)
#       + n_noise^2 Matern32(1, l_noise);  MarginalSparse(approx="VFE"), noise = sigma
CO2_PROG = (("periodic", "rbf"), ("rq",), ("rbf",), ("matern32",))
CO2_NAMES = ("log_n_per", "log_l_pdecay", "log_l_psmooth", "log_n_med", "log_l_med", "log_alpha", "log_n_trend", "log_l_trend",
             "log_n_noise", "log_l_noise", "sigma_log__")
CO2_PRIOR_SD = (3.0, 0.1, 1.0, 3.0, 3.0, 0.1, 3.0, 1.0, 3.0, 1.0)      # Normal(0, sd) on the ten logs; sigma ~ HalfNormal(1)


_CO2_SD = {}


def _co2_sd(device):
    """prior scales as a device constant (created once per device: no host-to-device copy inside a captured evaluation)"""
    key = str(device)
    if key not in _CO2_SD:
        _CO2_SD[key] = torch.tensor(CO2_PRIOR_SD, dtype=torch.float64, device=device)
    return _CO2_SD[key]


def co2_theta_from_x(x):
    """x[C, 11] (CO2_NAMES) -> theta[C, 12] of CO2_PROG in one input dimension: [a_per, ell_per = 2 l_psmooth (pymc3's Periodic is
    exp(-sin^2 / (2 ls^2))), period = 1, l_pdecay | a_med, l_med, alpha | a_trend, l_trend | a_noise, l_noise | sigma^2], and the
    Jacobian factors d theta / d x of the entries that move (the period is a constant)."""
    e = torch.exp(x)
    one = torch.ones_like(e[:, 0])
    theta = torch.stack([e[:, 0] ** 2, 2.0 * e[:, 2], one, e[:, 1], e[:, 3] ** 2, e[:, 4], e[:, 5], e[:, 6] ** 2, e[:, 7],
                         e[:, 8] ** 2, e[:, 9], e[:, 10] ** 2], dim=1)
    return theta


def co2_logp_dlogp(x, X, y, Z, jitter_policy="pymc3", engine=None, group=False):
    """pymc3 log-posterior of the reference's CO2 model and its gradient, batched over the rows of x[C, 11] (CO2_NAMES): the collapsed
    VFE bound with the composite kernel + Normal priors on the ten log-parameters + HalfNormal(1) on sigma with its log-transform
    Jacobian.  Rows whose factorisation fails get logp = -inf and a zero gradient."""
    if X.shape[1] != 1:
        raise ValueError("the CO2 model is one-dimensional (time)")
    eng = engine or Engine.get(X.device, CO2_PROG)
    x = x.to(device=eng.device, dtype=torch.float64)
    if x.dim() == 1:
        x = x.unsqueeze(0)
    theta = co2_theta_from_x(x)
    out = eng.sgpr_eval(X, y, Z, theta, jitter_policy=jitter_policy, need_grad=True, group=group, raise_on_fail=False)
    g = out["grad"]                                       # [C, 11 kernel parameters + s2 + dZ]
    th = theta
    # chain rule to x: amplitudes a = e^{2x}: 2 a dF/da ; lengths / alpha v = e^x: v dF/dv ; ell_per = 2 e^x: ell_per dF/d ell_per
    dx = torch.stack([2.0 * th[:, 0] * g[:, 0], th[:, 3] * g[:, 3], th[:, 1] * g[:, 1], 2.0 * th[:, 4] * g[:, 4], th[:, 5] * g[:, 5],
                      th[:, 6] * g[:, 6], 2.0 * th[:, 7] * g[:, 7], th[:, 8] * g[:, 8], 2.0 * th[:, 9] * g[:, 9], th[:, 10] * g[:, 10],
                      2.0 * th[:, 11] * g[:, 11]], dim=1)
    sd = _co2_sd(x.device)
    lp = out["bound"] + (-0.5 * (x[:, :10] / sd) ** 2 - torch.log(sd) - 0.5 * LOG2PI).sum(1) \
        + (0.5 * math.log(2.0 / math.pi) - 0.5 * th[:, 11] + x[:, 10])
    dx = dx + torch.cat([-x[:, :10] / sd ** 2, (1.0 - th[:, 11]).unsqueeze(1)], dim=1)
    info = torch.as_tensor(out["info"]).to(device=x.device)
    bad = (info != 0) | (out["info_b"] != 0) | ~torch.isfinite(lp)
    lp = torch.where(bad, torch.full_like(lp, -float("inf")), lp)
    dx = torch.where(bad.unsqueeze(1), torch.zeros_like(dx), dx)
    return lp, dx


def _vfe_eval(x, X, y, Z, jitter_policy, eng, group, with_prior):
    """Shared core of the pymc3 targets: x[C, D+2] unconstrained, one Z for all rows.  Returns (lp[C], dx[C, D+2], dZ[C, M, D], bad[C])."""
    D = X.shape[1]
    M = Z.shape[0]
    ell = torch.exp(x[:, :D])
    sig_f = torch.exp(x[:, D])
    sig_n = torch.exp(x[:, D + 1])
    theta = torch.cat([ell, (sig_f ** 2).unsqueeze(1), (sig_n ** 2).unsqueeze(1)], dim=1)
    out = eng.sgpr_eval(X, y, Z, theta, jitter_policy=jitter_policy, need_grad=True, group=group, raise_on_fail=False)
    g = out["grad"]
    dx = torch.cat([g[:, :D] * ell, (g[:, D] * 2.0 * sig_f ** 2).unsqueeze(1), (g[:, D + 1] * 2.0 * sig_n ** 2).unsqueeze(1)], dim=1)
    dZ = g[:, D + 2:].reshape(-1, M, D)
    lp = out["bound"].clone()
    if with_prior:
        lp = lp + (torch.log(ell) - ell).sum(1) + (LOG2 - LOGPI - torch.log1p(sig_f ** 2)) + (LOG2 - LOGPI - torch.log1p(sig_n ** 2)) \
            + x.sum(1)
        dpr = torch.cat([(1.0 - ell), (-2.0 * sig_f ** 2 / (1.0 + sig_f ** 2)).unsqueeze(1),
                         (-2.0 * sig_n ** 2 / (1.0 + sig_n ** 2)).unsqueeze(1)], dim=1) + 1.0
        dx = dx + dpr
    bad = (out["info"].to(eng.device) != 0) | (out["info_b"] != 0) | ~torch.isfinite(lp)
    return lp, dx, dZ, bad


def sgpr_vfe_logp_dlogp(x, X, y, Z, jitter_policy="pymc3", engine=None, group=False, with_prior=True):
    """pymc3 log-posterior of models/bayesian_sgpr_hmc.py:60-71 and its gradient, batched over the rows of x.

    x[C, D+2] is pymc3's unconstrained point (ls_log__[D], sig_f_log__, sig_n_log__): ell = e^x, sf2 = (e^x_f)^2,
    s2 = (e^x_n)^2; priors Gamma(2,1) on ell_d, HalfCauchy(1) on sig_f, sig_n, plus the log-Jacobians.
    Rows whose Cholesky fails get logp = -inf and a zero gradient (the sampler rejects them).
    """
    eng = engine or Engine.get(X.device)
    x = x.to(device=eng.device, dtype=torch.float64)
    if x.dim() == 1:
        x = x.unsqueeze(0)
    # transforms, priors, Jacobians and the failure mask in two launches (csrc/nuts.cuh k_vfe_theta / k_vfe_logp) instead of ~55
    # elementwise torch calls (what _vfe_eval spells out): at the reference's sizes a leapfrog is launch-latency bound
    from ._lib import check, load
    lib, st = load(), torch.cuda.current_stream().cuda_stream
    x = x.contiguous()
    C, D = x.shape[0], X.shape[1]
    theta = torch.empty(C, D + 2, dtype=torch.float64, device=x.device)
    check(lib.ggp_vfe_theta(st, x.data_ptr(), C, D, theta.data_ptr()), "ggp_vfe_theta")
    out = eng.sgpr_eval(X, y, Z, theta, jitter_policy=jitter_policy, need_grad=True, group=group, raise_on_fail=False)
    g, bound = out["grad"].contiguous(), out["bound"].contiguous()
    info = torch.as_tensor(out["info"]).to(device=x.device, dtype=torch.int32).contiguous()
    info_b = out["info_b"].to(device=x.device, dtype=torch.int32).contiguous()
    lp = torch.empty(C, dtype=torch.float64, device=x.device)
    dx = torch.empty(C, D + 2, dtype=torch.float64, device=x.device)
    check(lib.ggp_vfe_logp(st, x.data_ptr(), bound.data_ptr(), g.data_ptr(), g.stride(0), info.data_ptr(), info_b.data_ptr(),
                           C, D, 1 if with_prior else 0, lp.data_ptr(), dx.data_ptr()), "ggp_vfe_logp")
    return lp, dx


def all_in_hmc_logp_dlogp(x, X, y, num_inducing, jitter_policy="pymc3", engine=None, group=False):
    """Log-posterior of models/all_in_HMC.py:47-60 (Rossi et al. baseline): theta AND the inducing inputs Z are sampled.

    x[C, D+2+M*D] = (ls_log__[D], sig_f_log__, sig_n_log__, Z[M*D] row-major); Z ~ Normal(0,1) elementwise (:57), untransformed.
    dF/dZ is the analytic inducing-input gradient of the same bound+gradient evaluation.  Every chain carries its
    own Z, so chains are evaluated one after another (the reference runs chains=1, :60)."""
    eng = engine or Engine.get(X.device)
    x = x.to(device=eng.device, dtype=torch.float64)
    if x.dim() == 1:
        x = x.unsqueeze(0)
    D = X.shape[1]
    M = int(num_inducing)
    assert x.shape[1] == D + 2 + M * D
    lps, gs = [], []
    for c in range(x.shape[0]):
        Zc = x[c, D + 2:].reshape(M, D).contiguous()
        lp, dx, dZ, bad = _vfe_eval(x[c:c + 1, :D + 2], X, y, Zc, jitter_policy, eng, group, True)
        lp = lp[0] + (-0.5 * Zc * Zc).sum() - 0.5 * M * D * math.log(2.0 * math.pi)
        g = torch.cat([dx[0], (dZ[0] - Zc).reshape(-1)])
        b = bad[0] | ~torch.isfinite(lp)
        lps.append(torch.where(b, torch.full_like(lp, -float("inf")), lp))
        gs.append(torch.where(b, torch.zeros_like(g), g))
    return torch.stack(lps), torch.stack(gs)


class SVGPElbo(torch.autograd.Function):
    """Whitened SVGP minibatch ELBO: drop-in for  output = self(x_batch); -mll(output, y_batch)  (models/svgp.py:104-106).
    Differentiable in Z, q_mean, q_chol (lower triangle), lengthscale, outputscale, noise (constrained values)."""

    @staticmethod
    def forward(ctx, xb, yb, Z, q_mean, q_chol, lengthscale, outputscale, noise, num_data, cfg=None):
        c = _cfg(cfg)
        eng = Engine.get(xb.device, c["kernel"], "fp64" if c["precision"] == "auto" else c["precision"], c["chunk_rows"])
        theta = torch.cat([lengthscale.reshape(-1), outputscale.reshape(-1), noise.reshape(-1)]).to(torch.float64)
        need = any(ctx.needs_input_grad[2:8])
        out = eng.svgp_eval(xb, yb, Z, q_mean, q_chol, theta, num_data=num_data, likelihood=c.get("likelihood", "gaussian"),
                            jitter_policy=c["jitter_policy"], need_grad=need)
        ctx.dims = (xb.shape[1], Z.shape[0])
        ctx.shapes = tuple(t.shape for t in (Z, q_mean, q_chol, lengthscale, outputscale, noise))
        ctx.dtypes = tuple(t.dtype for t in (Z, q_mean, q_chol, lengthscale, outputscale, noise))
        ctx.save_for_backward(out["grad"][0] if need else torch.empty(0, device=xb.device))
        return out["value"][0].to(xb.dtype)

    @staticmethod
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        D, M = ctx.dims
        sh, dt = ctx.shapes, ctx.dtypes
        gout = gout.to(torch.float64)
        o = D + 2
        pick = lambda i, t, k: (gout * t).reshape(sh[k]).to(dt[k]) if ctx.needs_input_grad[i] else None
        gZ = pick(2, g[o:o + M * D], 0)
        gm = pick(3, g[o + M * D:o + M * D + M], 1)
        gL = pick(4, g[o + M * D + M:], 2)
        gl = pick(5, g[:D], 3)
        go = pick(6, g[D], 4)
        gn = pick(7, g[D + 1], 5)
        return None, None, gZ, gm, gL, gl, go, gn, None, None


def svgp_elbo(xb, yb, Z, q_mean, q_chol, lengthscale, outputscale, noise, num_data, cfg=None):
    return SVGPElbo.apply(xb, yb, Z, q_mean, q_chol, lengthscale, outputscale, noise, num_data, cfg)


def sgpmc_logp_dlogp(v, raw, X, y, Z, likelihood="gaussian", jitter=1e-5, engine=None, with_priors=True):
    """gpflow SGPMC log_posterior_density and gradient, batched over chains (models/sgp_hmc.py:38-83).

    v [C, M] whitened inducing values, raw [C, D+2] softplus-unconstrained (ell[D], sf2, s2).  Returns (logp[C], d/dv[C,M],
    d/draw[C,D+2]).  The streamed N x M work (a = L^{-1}k(Z,x), mu = a^T v, var = k - |a|^2, likelihood, backward) of ALL C chains
    runs in one batched launch sequence on the GPU (every chain carries its own theta and its own v: qm_batched); priors
    Gamma(2,1) on the constrained values + softplus log-Jacobians are added here, vectorised over the chains."""
    import torch.nn.functional as Fnn
    eng = engine or Engine.get(X.device)
    dev = eng.device
    v = v.to(device=dev, dtype=torch.float64)
    raw = raw.to(device=dev, dtype=torch.float64)
    if v.dim() == 1:
        v, raw = v.unsqueeze(0), raw.unsqueeze(0)
    C, M = v.shape
    D = X.shape[1]
    pos = Fnn.softplus(raw)
    theta = pos.clone()
    npos = D + 2 if likelihood == "gaussian" else D + 1
    if likelihood != "gaussian":
        theta[:, D + 1] = 1.0  # unused by the Bernoulli likelihood
    out = eng.svgp_eval(X, y, Z, v.contiguous(), None, theta, likelihood=likelihood, jitter_policy=0.0, base_jitter=jitter,
                        data_jitter=0.0, lik_scale=1.0, kl_scale=0.0, need_grad=True, raise_on_fail=False)
    g = out["grad"]
    lp = out["value"] - 0.5 * (v * v).sum(1) - 0.5 * M * math.log(2.0 * math.pi)
    gv = g[:, D + 2 + M * D:D + 2 + M * D + M] - v
    gpos = g[:, :D + 2].clone()
    if likelihood != "gaussian":
        gpos[:, D + 1] = 0.0
    sig = torch.sigmoid(raw)
    graw = gpos * sig
    if with_priors:
        lp = lp + (torch.log(pos[:, :npos]) - pos[:, :npos]).sum(1) + Fnn.logsigmoid(raw[:, :npos]).sum(1)
        graw[:, :npos] = graw[:, :npos] + (1.0 / pos[:, :npos] - 1.0) * sig[:, :npos] + (1.0 - sig[:, :npos])
    bad = (out["info"].to(dev) != 0) | ~torch.isfinite(lp)
    lp = torch.where(bad, torch.full_like(lp, -float("inf")), lp)
    gv = torch.where(bad.unsqueeze(1), torch.zeros_like(gv), gv)
    graw = torch.where(bad.unsqueeze(1), torch.zeros_like(graw), graw)
    return lp, gv, graw
