"""Lock-step batched HMC over many chains (host driver; the per-leapfrog logp/dlogp is ONE batched GPU evaluation).

Replaces the samplers the reference calls:
  * pm.sample(n, tune, chains=1, step=pm.NUTS())             models/bayesian_sgpr_hmc.py:73-78   (target: sgpr_vfe_logp_dlogp)
  * tfp.mcmc.HamiltonianMonteCarlo(L=10, step 0.01) + SimpleStepSizeAdaptation   models/sgp_hmc.py:67-83 (target: sgpmc)
Two samplers: fixed-length HMC (`hmc_sample`; every chain takes the same number of leapfrogs, so C chains cost one launch sequence
per leapfrog, and the whole trajectory can be replayed as one CUDA graph) and multinomial NUTS with pymc3's defaults
(`nuts_sample`; lock-step tree doubling with per-chain masking).  Both use per-chain dual-averaging step sizes (target accept 0.8,
pymc3/Stan constants) and windowed diagonal mass adaptation.
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


class GraphedTrajectory:
    """L leapfrog steps of C lock-step chains replayed as ONE CUDA graph.

    At the reference's own sizes (co2: N=545, M=100) a bound+gradient evaluation is ~90 short kernels: the arithmetic is
    microseconds, the launch/host latency a millisecond.  Capturing the whole trajectory (L evaluations + the leapfrog updates)
    removes the host from the inner loop.  Needs a sync-free target: a fixed jitter policy (pymc3's 1e-6, gpflow's 1e-5), for which
    Engine.factor does no host read-back, and static shapes.  Inputs/outputs live in static buffers (copied in / cloned out)."""

    def __init__(self, logp_dlogp, C, P, n_leapfrog, device, dtype=torch.float64):
        self.f, self.L = logp_dlogp, int(n_leapfrog)
        z = lambda *sh: torch.zeros(*sh, dtype=dtype, device=device)
        self.x, self.p, self.g, self.eps, self.inv_mass = z(C, P), z(C, P), z(C, P), z(C), torch.ones(C, P, dtype=dtype, device=device)
        self.graph, self.out = None, None

    def _run(self):
        xn, pn, gn, e = self.x, self.p, self.g, self.eps.unsqueeze(1)
        lpn = None
        for _ in range(self.L):
            pn = pn + 0.5 * e * gn
            xn = xn + e * pn * self.inv_mass
            lpn, gn = self.f(xn)
            pn = pn + 0.5 * e * gn
        return xn, pn, lpn, gn

    def capture(self, x, g):
        """Warm up on a side stream at a valid point (x, g), then capture."""
        self.x.copy_(x); self.g.copy_(g); self.p.zero_(); self.eps.fill_(1e-3)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                self._run()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run()
        return self

    def __call__(self, x, p, g, eps, inv_mass):
        self.x.copy_(x); self.p.copy_(p); self.g.copy_(g); self.eps.copy_(eps); self.inv_mass.copy_(inv_mass)
        self.graph.replay()
        return tuple(t.clone() for t in self.out)


class GraphedLogp:
    """One logp/dlogp evaluation of C lock-step chains replayed as ONE CUDA graph (static input / output buffers).  NUTS cannot
    capture a whole trajectory (the tree depth is data dependent) but every leapfrog is the same launch sequence: replaying it removes
    the host from the ~100 short kernels of a small-problem evaluation.  Needs a sync-free target (fixed jitter policy)."""

    def __init__(self, logp_dlogp, x):
        self.x = x.clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                logp_dlogp(self.x)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = logp_dlogp(self.x)

    def __call__(self, x):
        self.x.copy_(x)
        self.graph.replay()
        return self.out[0].clone(), self.out[1].clone()


def hmc_sample(logp_dlogp, x0, n_samples, tune=500, n_leapfrog=10, step_size=0.01, target_accept=0.8, adapt_mass=True,
               adaptation="dual_averaging", num_adaptation_steps=None, generator=None, progress=None, cuda_graph=False):
    """Run C chains in lock-step.  logp_dlogp(x[C,P]) -> (logp[C], grad[C,P]) (rows with logp=-inf are rejected).

    Returns dict(samples[n_samples, C, P], logp[n_samples, C], accept_rate[C], step_size[C], n_evals, inv_mass[C,P]).
    adaptation="dual_averaging" (pymc3-like) or "simple" (tfp SimpleStepSizeAdaptation: multiplicative, first
    `num_adaptation_steps` iterations only; models/sgp_hmc.py:71-73)."""
    x = x0.clone()
    C, P = x.shape
    dev, dt = x.device, x.dtype
    lp, g = logp_dlogp(x)
    n_evals = 1
    traj = GraphedTrajectory(logp_dlogp, C, P, n_leapfrog, dev, dt).capture(x, g) if (cuda_graph and x.is_cuda) else None
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
        if traj is not None:
            xn, pn, lpn, gn = traj(x, p, g, eps, inv_mass)
            n_evals += n_leapfrog
        else:
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


def _ckpt_range(n):
    """Checkpoint slots of the iterative U-turn scheme for leaf n of a subtree built left-to-right: every balanced
    sub-subtree that ENDS at an odd leaf n starts at an even leaf whose (momentum, running sum) was saved in slot
    popcount(start >> 1).  Returns (idx_min, idx_max): save into idx_max when n is even, check idx_max..idx_min when n is odd."""
    idx_max = bin(n >> 1).count("1")
    t, m = 0, n
    while m & 1:
        t, m = t + 1, m >> 1
    return idx_max - t + 1, idx_max


class RunningDiagMass:
    """pymc3's QuadPotentialDiagAdapt (what init='jitter+adapt_diag' builds for pm.NUTS(), models/bayesian_sgpr_hmc.py:73-78): the
    diagonal inverse metric is the running variance of the tuning draws, refreshed after EVERY tuning step from a foreground
    estimator that starts at (mean = start point, variance = 1, weight 10); every `window` draws the foreground is replaced by the
    background estimator (which only saw the last window) and a fresh background starts.  Works for any tune length (25 / 100 in
    models/bayesian_sgpr_hmc.py:141-146).  All state is [C, P]: one estimator per lock-step chain."""

    def __init__(self, x0, window=101, initial_weight=10):
        self.window, self.n = window, 0
        self.fg = [int(initial_weight), x0.clone(), torch.ones_like(x0) * initial_weight]   # n_samples, mean, raw_var
        self.bg = [0, torch.zeros_like(x0), torch.zeros_like(x0)]

    @staticmethod
    def _add(st, x):
        st[0] += 1
        old = x - st[1]
        st[1] = st[1] + old / st[0]
        st[2] = st[2] + old * (x - st[1])

    def update(self, x):
        self._add(self.fg, x)
        self._add(self.bg, x)
        var = self.fg[2] / self.fg[0]
        self.n += 1
        if self.n % self.window == 0:
            self.fg, self.bg = self.bg, [0, torch.zeros_like(x), torch.zeros_like(x)]
        return var


def nuts_sample(logp_dlogp, x0, n_samples, tune=500, target_accept=0.8, max_treedepth=10, step_size=None, adapt_mass=True,
                max_energy_error=1000.0, generator=None, progress=None, mass_adaptation="pymc3", cuda_graph=False, native=None):
    """No-U-turn sampler, C chains in lock-step, with the defaults of the sampler the reference calls: `pm.sample(n, tune=tune,
    chains=1)` with `pm.NUTS()` (models/bayesian_sgpr_hmc.py:73-78, models/all_in_HMC.py:60): multinomial NUTS, uniform progressive
    sampling inside a subtree and biased progressive sampling between the old tree and the new subtree, U-turn test
    p_sum . v_edge <= 0 at both edges of every balanced subtree, divergence at |energy change| > 1000, max_treedepth 10,
    step size by dual averaging to target_accept 0.8 from 0.25 / P^(1/4), statistic = mean over tree leaves of min(1, exp(-dE)).
    The mass matrix is diagonal; mass_adaptation="pymc3" (default) is pymc3's running estimator (RunningDiagMass: refreshed every
    tuning step, foreground / background windows of 101 draws, no restart of the step-size adaptation), "windows" the Stan-like
    three-window scheme hmc_sample uses.

    Batching: one tree doubling = 2^j leapfrogs = 2^j BATCHED logp/dlogp calls shared by every chain; a chain whose tree has
    stopped is masked out (its rows are still evaluated, which is what lock-step costs).  The U-turn checks inside a subtree use
    the iterative checkpoint scheme (no recursion), so the leaf index - and hence the control flow - is identical for all chains.
    Random numbers: one normal draw [C, P] per transition and one uniform draw [2 + 2^j, C] per doubling j (row 0 direction, row 1
    merge, rows 2.. one per leaf), so the stream does not depend on who does the bookkeeping.

    native (default: x0 is on a GPU): the tree bookkeeping between two evaluations runs in the CUDA kernels of csrc/nuts.cuh
    (NativeNutsTree below: one launch per leaf, in the same CUDA graph as the evaluation when cuda_graph=True) instead of ~60
    elementwise torch calls per leaf; same algorithm, same random numbers, positions bit-identical per leapfrog.  native=False is
    the torch implementation below (any device; the CPU tests and the GPU cross-check use it).
    Returns the hmc_sample dict plus tree_depth[n_samples, C], n_leapfrog[n_samples, C], diverging[n_samples, C]."""
    if native is None:
        native = x0.is_cuda
    if native:
        return _nuts_sample_native(logp_dlogp, x0, n_samples, tune, target_accept, max_treedepth, step_size, adapt_mass,
                                   max_energy_error, generator, progress, mass_adaptation, cuda_graph)
    x = x0.clone()
    C, P = x.shape
    dev, dt = x.device, x.dtype
    U = lambda *sh: torch.rand(*sh, dtype=dt, device=dev, generator=generator)
    if cuda_graph and x.is_cuda:
        logp_dlogp = GraphedLogp(logp_dlogp, x)     # every leapfrog evaluation = one graph replay (bit-identical to the eager call)
    lp, g = logp_dlogp(x)
    n_evals = 1
    eps0 = float(step_size) if step_size is not None else 0.25 / P ** 0.25
    eps = torch.full((C,), eps0, dtype=dt, device=dev)
    inv_mass = torch.ones(C, P, dtype=dt, device=dev)
    da = DualAveraging(eps, target_accept)
    samples = torch.empty(n_samples, C, P, dtype=dt, device=dev)
    lps = torch.empty(n_samples, C, dtype=dt, device=dev)
    depths = torch.zeros(n_samples, C, dtype=torch.int32, device=dev)
    nleap = torch.zeros(n_samples, C, dtype=torch.int32, device=dev)
    divs = torch.zeros(n_samples, C, dtype=torch.bool, device=dev)
    acc_sum = torch.zeros(C, dtype=dt, device=dev)
    running = RunningDiagMass(x) if (adapt_mass and mass_adaptation == "pymc3") else None
    windows = sorted({int(tune * f) for f in (0.25, 0.5, 0.75)} - {0}) if adapt_mass and running is None and tune >= 40 else []
    win_start, buf = 0, []
    ninf = torch.full((C,), -float("inf"), dtype=dt, device=dev)
    K = max(max_treedepth, 1)

    def turning(p_left, p_right, p_sum):
        return ((p_sum * p_left * inv_mass).sum(1) <= 0) | ((p_sum * p_right * inv_mass).sum(1) <= 0)

    n_evals_tune = n_evals
    for it in range(tune + n_samples):
        if it == tune:
            n_evals_tune = n_evals
        p0 = torch.randn(C, P, dtype=dt, device=dev, generator=generator) / torch.sqrt(inv_mass)
        e0 = -lp + 0.5 * (p0 * p0 * inv_mass).sum(1)
        xl, pl, gl = x, p0, g
        xr, pr, gr = x, p0, g
        x_prop, lp_prop, g_prop = x, lp, g
        log_w = torch.zeros(C, dtype=dt, device=dev)
        p_sum = p0.clone()
        sum_acc = torch.zeros(C, dtype=dt, device=dev)
        n_leaf = torch.zeros(C, dtype=dt, device=dev)
        depth = torch.zeros(C, dtype=torch.int32, device=dev)
        diverged = torch.zeros(C, dtype=torch.bool, device=dev)
        active = torch.ones(C, dtype=torch.bool, device=dev)
        for j in range(max_treedepth):
            if not bool(active.any()):
                break
            u = U(2 + 2 ** j, C)
            right = u[0] < 0.5
            sgn = torch.where(right, torch.ones_like(eps), -torch.ones_like(eps))
            e = (eps * sgn).unsqueeze(1)
            r1 = right.unsqueeze(1)
            xe, pe, ge = torch.where(r1, xr, xl), torch.where(r1, pr, pl), torch.where(r1, gr, gl)
            s_log_w, s_p_sum = ninf.clone(), torch.zeros(C, P, dtype=dt, device=dev)
            s_x, s_lp, s_g = xe, lp, ge
            s_turn = torch.zeros(C, dtype=torch.bool, device=dev)
            s_div = torch.zeros(C, dtype=torch.bool, device=dev)
            building = active.clone()
            p_ck = torch.zeros(K, C, P, dtype=dt, device=dev)
            ps_ck = torch.zeros(K, C, P, dtype=dt, device=dev)
            for n in range(2 ** j):
                pn = pe + 0.5 * e * ge
                xn = xe + e * pn * inv_mass
                lpn, gn = logp_dlogp(xn)
                n_evals += 1
                pn = pn + 0.5 * e * gn
                en = -lpn + 0.5 * (pn * pn * inv_mass).sum(1)
                dlt = e0 - en                                              # log weight of the leaf
                bad = ~torch.isfinite(dlt) | (dlt.abs() > max_energy_error)
                s_div = s_div | (building & bad)
                dlt = torch.where(bad, ninf, dlt)
                good = building & ~bad
                sum_acc = sum_acc + torch.where(building, torch.exp(dlt.clamp(max=0.0)), torch.zeros_like(dlt))
                n_leaf = n_leaf + building.to(dt)
                new_w = torch.logaddexp(s_log_w, dlt)
                take = good & (torch.log(u[2 + n]) < dlt - new_w)
                t1 = take.unsqueeze(1)
                s_x, s_lp, s_g = torch.where(t1, xn, s_x), torch.where(take, lpn, s_lp), torch.where(t1, gn, s_g)
                s_log_w = torch.where(good, new_w, s_log_w)
                g1 = good.unsqueeze(1)
                s_p_sum = torch.where(g1, s_p_sum + pn, s_p_sum)
                xe, pe, ge = torch.where(g1, xn, xe), torch.where(g1, pn, pe), torch.where(g1, gn, ge)
                building = good
                lo, hi = _ckpt_range(n)
                if n % 2 == 0:
                    p_ck[hi], ps_ck[hi] = pe, s_p_sum
                else:
                    tn = torch.zeros(C, dtype=torch.bool, device=dev)
                    for i in range(hi, lo - 1, -1):
                        tn = tn | turning(p_ck[i], pe, s_p_sum - ps_ck[i] + p_ck[i])
                    s_turn = s_turn | (building & tn)
                    building = building & ~tn
            diverged = diverged | (active & s_div)
            ok = active & ~s_turn & ~s_div
            take = ok & (torch.log(u[1]) < s_log_w - log_w)
            t1 = take.unsqueeze(1)
            x_prop, lp_prop, g_prop = torch.where(t1, s_x, x_prop), torch.where(take, s_lp, lp_prop), torch.where(t1, s_g, g_prop)
            log_w = torch.where(ok, torch.logaddexp(log_w, s_log_w), log_w)
            o1 = ok.unsqueeze(1)
            p_sum = torch.where(o1, p_sum + s_p_sum, p_sum)
            okr, okl = (ok & right).unsqueeze(1), (ok & ~right).unsqueeze(1)
            xr, pr, gr = torch.where(okr, xe, xr), torch.where(okr, pe, pr), torch.where(okr, ge, gr)
            xl, pl, gl = torch.where(okl, xe, xl), torch.where(okl, pe, pl), torch.where(okl, ge, gl)
            depth = depth + ok.to(torch.int32)
            active = ok & ~turning(pl, pr, p_sum)
        x, lp, g = x_prop, lp_prop, g_prop
        acc_prob = sum_acc / n_leaf.clamp(min=1.0)
        if it < tune:
            eps = da.update(acc_prob)
            if it == tune - 1:
                eps = da.final()
            if running is not None:
                inv_mass = running.update(x)
            if windows:
                buf.append(x.clone())
                if it + 1 in windows:
                    S = torch.stack(buf[win_start:])
                    if S.shape[0] >= 10:
                        var = S.var(0, unbiased=True)
                        nn_ = S.shape[0]
                        inv_mass = (nn_ / (nn_ + 5.0)) * var + 1e-3 * (5.0 / (nn_ + 5.0))
                        da = DualAveraging(eps, target_accept)
                    win_start = len(buf)
        else:
            k = it - tune
            samples[k], lps[k], depths[k], nleap[k], divs[k] = x, lp, depth, n_leaf.to(torch.int32), diverged
            acc_sum += acc_prob
        if progress is not None:
            progress(it)
    return dict(samples=samples, logp=lps, accept_rate=acc_sum / max(n_samples, 1), step_size=eps, n_evals=n_evals,
                n_evals_sampling=n_evals - n_evals_tune, inv_mass=inv_mass, tree_depth=depths, n_leapfrog=nleap, diverging=divs)


class NativeNutsTree:
    """Device-resident NUTS tree of C lock-step chains (csrc/nuts.cuh through the C ABI: ggp_nuts_*).  Owns the state buffers of
    ggp_nuts_state; `leaf()` = one logp/dlogp evaluation at x_eval followed by ONE bookkeeping launch, replayed as a single CUDA
    graph when cuda_graph=True (the leaf index is a device counter, so every leaf is the same launch sequence)."""

    def __init__(self, logp_dlogp, x0, n_samples, max_treedepth, max_energy_error, cuda_graph):
        import ctypes
        from . import _lib
        if not x0.is_cuda:
            raise RuntimeError("NativeNutsTree needs CUDA tensors (there is no CPU build of csrc/nuts.cuh); use native=False")
        self.lib, self._check, self._byref = _lib.load(), _lib.check, ctypes.byref
        self.f = logp_dlogp
        C, P = x0.shape
        K = max(int(max_treedepth), 1)
        dev = x0.device
        zd = lambda *sh: torch.zeros(*sh, dtype=torch.float64, device=dev)
        zi = lambda *sh: torch.zeros(*sh, dtype=torch.int32, device=dev)
        t = {n: zd(C, P) for n in _lib.NUTS_DOUBLE_CP}
        t.update({n: zd(C) for n in _lib.NUTS_DOUBLE_C})
        t.update({n: zi(C) for n in _lib.NUTS_INT_C})
        ns = max(int(n_samples), 1)
        t.update(p_ck=zd(K, C, P), ps_ck=zd(K, C, P), any_active=zi(1), u=zd(2 + 2 ** K, C), samples=zd(ns, C, P), lps=zd(ns, C),
                 depths=zi(ns, C), nleaps=zi(ns, C), divs=zi(ns, C))
        t["inv_mass"].fill_(1.0)
        self.t = t
        self.st = _lib.GgpNutsState(C=C, P=P, K=K, max_energy_error=float(max_energy_error))
        for n in _lib._NUTS_PTRS:
            setattr(self.st, n, t[n].data_ptr())
        self.graph, self._keep = None, None
        lp, g = logp_dlogp(x0)
        t["x"].copy_(x0); t["lp"].copy_(lp); t["g"].copy_(g); t["x_eval"].copy_(x0)
        self.n_evals = 1
        if cuda_graph:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):
                    logp_dlogp(t["x_eval"])
            cur.wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._eval_and_leaf()

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def _eval_and_leaf(self):
        lp, g = self.f(self.t["x_eval"])
        lp, g = lp.to(torch.float64).contiguous(), g.to(torch.float64).contiguous()
        self._keep = (lp, g)                      # graph mode: the captured outputs; eager: alive until the next evaluation
        self.st.lp_eval, self.st.g_eval = lp.data_ptr(), g.data_ptr()
        self._check(self.lib.ggp_nuts_leaf(self._stream(), self._byref(self.st)), "ggp_nuts_leaf")

    def begin(self, z):
        self._check(self.lib.ggp_nuts_begin(self._stream(), self._byref(self.st), z.data_ptr()), "ggp_nuts_begin")

    def doubling(self, j, generator):
        """Grow the tree by 2^j leaves in a random direction; True while some chain keeps doubling (the one host read)."""
        t = self.t
        torch.rand(2 + 2 ** j, t["u"].shape[1], dtype=torch.float64, device=t["u"].device, generator=generator, out=t["u"][:2 + 2 ** j])
        self._check(self.lib.ggp_nuts_subtree_begin(self._stream(), self._byref(self.st)), "ggp_nuts_subtree_begin")
        for _ in range(2 ** j):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._eval_and_leaf()
        self.n_evals += 2 ** j
        self._check(self.lib.ggp_nuts_subtree_end(self._stream(), self._byref(self.st)), "ggp_nuts_subtree_end")
        return bool(t["any_active"].item())

    def end(self, k):
        self._check(self.lib.ggp_nuts_end(self._stream(), self._byref(self.st), int(k)), "ggp_nuts_end")


def _nuts_sample_native(logp_dlogp, x0, n_samples, tune, target_accept, max_treedepth, step_size, adapt_mass, max_energy_error,
                        generator, progress, mass_adaptation, cuda_graph):
    """nuts_sample with the tree in csrc/nuts.cuh; adaptation (once per transition) stays in torch."""
    C, P = x0.shape
    dev, dt = x0.device, x0.dtype
    tree = NativeNutsTree(logp_dlogp, x0.to(torch.float64), n_samples, max_treedepth, max_energy_error, cuda_graph)
    t = tree.t
    eps0 = float(step_size) if step_size is not None else 0.25 / P ** 0.25
    t["eps"].fill_(eps0)
    da = DualAveraging(t["eps"].clone(), target_accept)
    running = RunningDiagMass(t["x"].clone()) if (adapt_mass and mass_adaptation == "pymc3") else None
    windows = sorted({int(tune * f) for f in (0.25, 0.5, 0.75)} - {0}) if adapt_mass and running is None and tune >= 40 else []
    win_start, buf = 0, []
    acc_sum = torch.zeros(C, dtype=torch.float64, device=dev)
    n_evals_tune = tree.n_evals
    for it in range(tune + n_samples):
        if it == tune:
            n_evals_tune = tree.n_evals
        tree.begin(torch.randn(C, P, dtype=torch.float64, device=dev, generator=generator))
        for j in range(max_treedepth):
            if not tree.doubling(j, generator):
                break
        tree.end(it - tune if it >= tune else -1)
        if it < tune:
            eps = da.update(t["acc_prob"])
            if it == tune - 1:
                eps = da.final()
            t["eps"].copy_(eps)
            if running is not None:
                t["inv_mass"].copy_(running.update(t["x"]))
            if windows:
                buf.append(t["x"].clone())
                if it + 1 in windows:
                    S = torch.stack(buf[win_start:])
                    if S.shape[0] >= 10:
                        var = S.var(0, unbiased=True)
                        nn_ = S.shape[0]
                        t["inv_mass"].copy_((nn_ / (nn_ + 5.0)) * var + 1e-3 * (5.0 / (nn_ + 5.0)))
                        da = DualAveraging(t["eps"].clone(), target_accept)
                    win_start = len(buf)
        else:
            acc_sum += t["acc_prob"]
        if progress is not None:
            progress(it)
    return dict(samples=t["samples"][:n_samples].to(dt), logp=t["lps"][:n_samples].to(dt), accept_rate=acc_sum / max(n_samples, 1),
                step_size=t["eps"].clone(), n_evals=tree.n_evals, n_evals_sampling=tree.n_evals - n_evals_tune, inv_mass=t["inv_mass"].clone(), tree_depth=t["depths"][:n_samples],
                n_leapfrog=t["nleaps"][:n_samples], diverging=t["divs"][:n_samples].bool(), native=True)


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


def sample_hyper(X, y, Z, n_samples, tune, chains=1, n_leapfrog=10, step_size=None, engine=None, generator=None, seed_jitter=True,
                 cuda_graph=None, sampler="nuts", max_treedepth=10, target_accept=0.8):
    """Sampling over theta = (ls, sig_f, sig_n) on the collapsed VFE bound with pymc3's priors and transforms
    (models/bayesian_sgpr_hmc.py:58-80).  Start = prior test value (Gamma mean 2, HalfCauchy beta 1) + U(-1,1) jitter
    in unconstrained space (pymc3 init='jitter+adapt_diag').  sampler="nuts" (default) is pm.NUTS() as the reference calls it
    (:73-78); sampler="hmc" is the fixed-length lock-step sampler (step_size default 0.02; the CUDA-graph benchmark uses it).
    cuda_graph=None replays the evaluation (and the NUTS bookkeeping launch) as a CUDA graph when the problem is small enough for
    launch latency to matter (N * M <= 2^22, the reference's own data sets); larger problems run eagerly.
    Returns (list of HyperTrace per chain, raw result)."""
    import time
    if cuda_graph is None:
        cuda_graph = bool(X.is_cuda and X.shape[0] * Z.shape[0] <= 2 ** 22)
    from .functions import sgpr_vfe_logp_dlogp
    D = X.shape[1]
    dev = X.device
    x0 = torch.zeros(chains, D + 2, dtype=torch.float64, device=dev)
    x0[:, :D] = math.log(2.0)
    if seed_jitter:
        x0 = x0 + (torch.rand(chains, D + 2, dtype=torch.float64, device=dev, generator=generator) * 2.0 - 1.0)
    f = lambda xx: sgpr_vfe_logp_dlogp(xx, X, y, Z, engine=engine)
    t0 = time.perf_counter()
    if sampler == "nuts":   # pm.NUTS() defaults (models/bayesian_sgpr_hmc.py:73-78)
        res = nuts_sample(f, x0, n_samples, tune=tune, max_treedepth=max_treedepth, target_accept=target_accept, step_size=step_size,
                          generator=generator, cuda_graph=cuda_graph)
    elif sampler == "hmc":
        res = hmc_sample(f, x0, n_samples, tune=tune, n_leapfrog=n_leapfrog, step_size=0.02 if step_size is None else step_size,
                         target_accept=target_accept, generator=generator, cuda_graph=cuda_graph)
    else:
        raise ValueError(f"unknown sampler {sampler!r}: 'nuts' or 'hmc'")
    dt = time.perf_counter() - t0
    res["seconds"] = dt
    traces = [HyperTrace(res["samples"][:, c], res["step_size"][c], dt) for c in range(chains)]
    return traces, res
