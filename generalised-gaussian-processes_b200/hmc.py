"""Lock-step batched HMC over many chains (host driver; the per-leapfrog logp/dlogp is ONE batched GPU evaluation).

Replaces the samplers the reference calls:
  * pm.sample(n, tune, chains=1, step=pm.NUTS())             models/bayesian_sgpr_hmc.py:73-78   (target: sgpr_vfe_logp_dlogp)
  * tfp.mcmc.HamiltonianMonteCarlo(L=10, step 0.01) + SimpleStepSizeAdaptation   models/sgp_hmc.py:67-83 (target: sgpmc)
This round ships fixed-length HMC (it batches trivially: every chain takes the same number of leapfrogs, so C chains cost one
launch sequence per leapfrog) with per-chain dual-averaging step sizes (target accept 0.8, pymc3/Stan constants) and windowed
diagonal mass adaptation.  NUTS with per-chain tree masking is the next row; DESIGN.md says so.
"""
import math

import torch


class DualAveraging:
    """Nesterov dual averaging of log step size (Hoffman & Gelman 2014; pymc3 step_size adaptation constants)."""

    def __init__(self, eps0, target=0.8, gamma=0.05, t0=10.0, kappa=0.75):
        self.mu = torch.log(10.0 * eps0)
        self.target, self.gamma, self.t0, self.kappa = target, gamma, t0, kappa
        self.hbar = torch.zeros_like(eps0)
        self.log_eps_bar = torch.zeros_like(eps0)
        self.t = 0

    def update(self, accept_prob):
        self.t += 1
        w = 1.0 / (self.t + self.t0)
        self.hbar = (1.0 - w) * self.hbar + w * (self.target - accept_prob)
        log_eps = self.mu - math.sqrt(self.t) / self.gamma * self.hbar
        eta = self.t ** (-self.kappa)
        self.log_eps_bar = eta * log_eps + (1.0 - eta) * self.log_eps_bar
        return torch.exp(log_eps)

    def final(self):
        return torch.exp(self.log_eps_bar)


def hmc_sample(logp_dlogp, x0, n_samples, tune=500, n_leapfrog=10, step_size=0.01, target_accept=0.8, adapt_mass=True,
               adaptation="dual_averaging", num_adaptation_steps=None, generator=None, progress=None):
    """Run C chains in lock-step.  logp_dlogp(x[C,P]) -> (logp[C], grad[C,P]) (rows with logp=-inf are rejected).

    Returns dict(samples[n_samples, C, P], logp[n_samples, C], accept_rate[C], step_size[C], n_evals, inv_mass[C,P]).
    adaptation="dual_averaging" (pymc3-like) or "simple" (tfp SimpleStepSizeAdaptation: multiplicative, first
    `num_adaptation_steps` iterations only; models/sgp_hmc.py:71-73)."""
    x = x0.clone()
    C, P = x.shape
    dev, dt = x.device, x.dtype
    lp, g = logp_dlogp(x)
    n_evals = 1
    eps = torch.full((C,), float(step_size), dtype=dt, device=dev)
    inv_mass = torch.ones(C, P, dtype=dt, device=dev)          # diagonal inverse metric (= posterior variance estimate)
    da = DualAveraging(eps, target_accept)
    samples = torch.empty(n_samples, C, P, dtype=dt, device=dev)
    lps = torch.empty(n_samples, C, dtype=dt, device=dev)
    acc_sum = torch.zeros(C, dtype=dt, device=dev)
    windows = sorted({int(tune * f) for f in (0.25, 0.5, 0.75)} - {0}) if adapt_mass and tune >= 40 else []
    win_start, buf = 0, []
    for it in range(tune + n_samples):
        p = torch.randn(C, P, dtype=dt, device=dev, generator=generator) / torch.sqrt(inv_mass)
        h0 = -lp + 0.5 * (p * p * inv_mass).sum(1)
        xn, pn, lpn, gn = x, p, lp, g
        e = eps.unsqueeze(1)
        for _ in range(n_leapfrog):
            pn = pn + 0.5 * e * gn
            xn = xn + e * pn * inv_mass
            lpn, gn = logp_dlogp(xn)
            n_evals += 1
            pn = pn + 0.5 * e * gn
        h1 = -lpn + 0.5 * (pn * pn * inv_mass).sum(1)
        dh = h0 - h1
        acc_prob = torch.where(torch.isfinite(dh), torch.exp(dh.clamp(max=0.0)), torch.zeros_like(dh))
        u = torch.rand(C, dtype=dt, device=dev, generator=generator)
        acc = u < acc_prob
        x = torch.where(acc.unsqueeze(1), xn, x)
        g = torch.where(acc.unsqueeze(1), gn, g)
        lp = torch.where(acc, lpn, lp)
        if it < tune:
            if adaptation == "dual_averaging":
                eps = da.update(acc_prob)
                if it == tune - 1:
                    eps = da.final()
            elif num_adaptation_steps is None or it < num_adaptation_steps:
                eps = torch.where(acc_prob > target_accept, eps * 1.1, eps / 1.1)
            if windows:
                buf.append(x.clone())
                if it + 1 in windows:
                    S = torch.stack(buf[win_start:])
                    if S.shape[0] >= 10:
                        var = S.var(0, unbiased=True)
                        n = S.shape[0]
                        inv_mass = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0))   # Stan's shrinkage towards 1e-3
                        if adaptation == "dual_averaging":
                            da = DualAveraging(eps, target_accept)
                    win_start = len(buf)
        else:
            k = it - tune
            samples[k], lps[k] = x, lp
            acc_sum += acc_prob
        if progress is not None:
            progress(it)
    return dict(samples=samples, logp=lps, accept_rate=acc_sum / max(n_samples, 1), step_size=eps, n_evals=n_evals,
                inv_mass=inv_mass, n_leapfrog=n_leapfrog)


class HyperTrace:
    """Minimal pymc3-MultiTrace look-alike: len(trace), trace[i] -> {'ls', 'sig_f', 'sig_n'} (numpy, float64),
    trace.get_sampler_stats('step_size').  Built from unconstrained log-space draws x[n, D+2] of ONE chain."""

    def __init__(self, x, step_size, perf_seconds):
        self.x = x.detach().cpu()
        self.D = x.shape[1] - 2
        self._stats = {"step_size": [float(step_size)] * len(self), "perf_counter_diff": [perf_seconds / max(len(self), 1)] * len(self)}

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        v = torch.exp(self.x[i])
        return {"ls": v[:self.D].numpy(), "sig_f": float(v[self.D]), "sig_n": float(v[self.D + 1])}

    def get_sampler_stats(self, name):
        import numpy as np
        return np.asarray(self._stats[name])

    def thetas(self):
        """[n, D+2] constrained (ell, sf2=sig_f^2, s2=sig_n^2) for batched evaluation (update_model_to_hyper, :82-86)."""
        v = torch.exp(self.x)
        return torch.cat([v[:, :self.D], v[:, self.D:] ** 2], dim=1)


def sample_hyper(X, y, Z, n_samples, tune, chains=1, n_leapfrog=10, step_size=0.02, engine=None, generator=None, seed_jitter=True):
    """HMC over theta = (ls, sig_f, sig_n) on the collapsed VFE bound with pymc3's priors and transforms
    (models/bayesian_sgpr_hmc.py:58-80).  Start = prior test value (Gamma mean 2, HalfCauchy beta 1) + U(-1,1) jitter
    in unconstrained space (pymc3 init='jitter+adapt_diag').  Returns (list of HyperTrace per chain, raw result)."""
    import time
    from .functions import sgpr_vfe_logp_dlogp
    D = X.shape[1]
    dev = X.device
    x0 = torch.zeros(chains, D + 2, dtype=torch.float64, device=dev)
    x0[:, :D] = math.log(2.0)
    if seed_jitter:
        x0 = x0 + (torch.rand(chains, D + 2, dtype=torch.float64, device=dev, generator=generator) * 2.0 - 1.0)
    f = lambda xx: sgpr_vfe_logp_dlogp(xx, X, y, Z, engine=engine)
    t0 = time.perf_counter()
    res = hmc_sample(f, x0, n_samples, tune=tune, n_leapfrog=n_leapfrog, step_size=step_size, generator=generator)
    dt = time.perf_counter() - t0
    res["seconds"] = dt
    traces = [HyperTrace(res["samples"][:, c], res["step_size"][c], dt) for c in range(chains)]
    return traces, res
