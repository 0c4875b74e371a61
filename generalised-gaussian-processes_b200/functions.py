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

DEFAULT_CFG = dict(normalize="n", jitter_policy="gpytorch", precision="fp64", kernel="rbf", group=None, chunk_rows=0)


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
        eng = Engine.get(X.device, c["kernel"], c["precision"], c["chunk_rows"])
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


LOG2 = math.log(2.0)
LOGPI = math.log(math.pi)


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
    D = X.shape[1]
    ell = torch.exp(x[:, :D])
    sig_f = torch.exp(x[:, D])
    sig_n = torch.exp(x[:, D + 1])
    theta = torch.cat([ell, (sig_f ** 2).unsqueeze(1), (sig_n ** 2).unsqueeze(1)], dim=1)
    out = eng.sgpr_eval(X, y, Z, theta, jitter_policy=jitter_policy, need_grad=True, group=group, raise_on_fail=False)
    g = out["grad"]
    dx = torch.cat([g[:, :D] * ell, (g[:, D] * 2.0 * sig_f ** 2).unsqueeze(1), (g[:, D + 1] * 2.0 * sig_n ** 2).unsqueeze(1)], dim=1)
    lp = out["bound"].clone()
    if with_prior:
        lp = lp + (torch.log(ell) - ell).sum(1) + (LOG2 - LOGPI - torch.log1p(sig_f ** 2)) + (LOG2 - LOGPI - torch.log1p(sig_n ** 2)) \
            + x.sum(1)
        dpr = torch.cat([(1.0 - ell), (-2.0 * sig_f ** 2 / (1.0 + sig_f ** 2)).unsqueeze(1),
                         (-2.0 * sig_n ** 2 / (1.0 + sig_n ** 2)).unsqueeze(1)], dim=1) + 1.0
        dx = dx + dpr
    bad = (out["info"].to(eng.device) != 0) | (out["info_b"] != 0) | ~torch.isfinite(lp)
    lp = torch.where(bad, torch.full_like(lp, -float("inf")), lp)
    dx = torch.where(bad.unsqueeze(1), torch.zeros_like(dx), dx)
    return lp, dx
