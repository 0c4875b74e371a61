#!/usr/bin/env python
"""Headline benchmark: SGPR collapsed bound + gradient evaluations/s at N=1e6, M=1024, D=8, FP64
(BASELINE.json configs[3], row-sharded over --gpus ranks with one all-reduce per pass).

  python bench.py --gpus 1 --steps K --warmup W              this repo's CUDA path
  python bench.py --impl reference ...                       the CPU restatement of the reference path (oracle/), all host
                                                             threads, FULL N per step (a time budget bounds the step count), rank 0
One JSON line on stdout (rank 0).  A "step" is one bound+gradient evaluation over all N rows.

Extra legs on the same line (all time-bounded): `parity_at_headline` (this run's outputs against the committed long-double
evaluation of the full configuration and against the float64 oracle run on the host), `variants` (with-replacement Z,
gpytorch-init theta, streaming mode), `hmc` (configs[1]: fixed-length HMC and pm.NUTS defaults), `config5_hmc` (configs[4]:
chain-batched SGPMC, chains sharded over the ranks), `svgp` (configs[2]) and `config1` (configs[0]); each with the CPU oracle timed
beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FULL, M_IND, D_IN = 1_000_000, 1024, 8
METRIC = "sgpr_bound_grad_evals_per_s"
WORKLOAD = "configs[3]: synthetic large_scale_regression N=1e6 D=8 M=1024 FP64 SGPR bound+grad, rows sharded over ranks"


def workload_config(N):
    """The `config` both arms print (the same dict: the driver compares them)."""
    return {"workload": WORKLOAD, "N": N, "M": M_IND, "D": D_IN, "theta": "trained-like (ell=sqrt(D), sf2=1, s2=0.1)",
            "Z": "1024 distinct training rows (without replacement)", "jitter_policy": "gpytorch",
            "l2": "256 MB buffer written between steps (inside the timed region)"}


def env_int(k, d):
    return int(os.environ.get(k, d))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------------
# CPU side (oracle/): the only places bench.py executes the checker -- the cpu_baseline legs and --impl reference
# ---------------------------------------------------------------------------------------------------------------------------
def cpu_oracle_full_eval(c, th, threads):
    """ONE float64 oracle bound+gradient evaluation on ALL rows of the headline configuration (chunked; ~30 s on 16 threads).
    Returns (seconds, F, gradient dict, jitter) -- not divided by N."""
    import torch
    from oracle import sgpr as osgpr
    torch.set_num_threads(threads)
    X, y, Z, tht = [torch.tensor(c[k]) for k in ("X", "y", "Z")] + [torch.tensor(th)]
    t0 = time.perf_counter()
    F, g, jit = osgpr.sgpr_bound_and_grads_chunked(X, y, Z, tht[:D_IN], tht[D_IN], tht[D_IN + 1], jitter_policy="gpytorch",
                                                   normalize="none", chunk=65536)
    return time.perf_counter() - t0, float(F), g, jit


def grad_blocks_relerr(grad, F, ref_F, ref_g, M, D):
    """max-norm relative error per gradient block of grad[d+2+m*d] / bound F against a reference (dict of arrays)."""
    import numpy as np
    g = np.asarray(grad, dtype=np.float64).reshape(-1)
    arr = lambda a: np.asarray(a, dtype=np.float64)
    rel = lambda a, b: float(np.abs(arr(a) - arr(b)).max() / np.abs(arr(b)).max())
    return {"bound": abs(F - ref_F) / abs(ref_F), "ell": rel(g[:D], ref_g["ell"]), "sf2": rel(g[D], ref_g["sf2"]),
            "s2": rel(g[D + 1], ref_g["s2"]), "Z": rel(g[D + 2:].reshape(M, D), ref_g["Z"])}


def run_reference(args, rank):
    """CPU arm: the float64 oracle port on all host threads, FULL N per step.  A wall-clock budget (GGP_REF_BUDGET_S, default 240 s)
    bounds the number of steps actually executed; the line reports the counts it ran."""
    if rank != 0:
        return
    import ggp_b200.synthetic as syn
    threads = os.cpu_count() or 1
    budget = float(os.environ.get("GGP_REF_BUDGET_S", 240.0))
    c = syn.config4_large(N=args.rows, D=D_IN, M=M_IND)
    th = syn.theta_trained_like(D_IN)
    t_start = time.perf_counter()
    warm = min(args.warmup, 1)
    times = []
    for i in range(warm + args.steps):
        dt, _, _, _ = cpu_oracle_full_eval(c, th, threads)
        if i >= warm:
            times.append(dt)
        if i >= warm and time.perf_counter() - t_start + dt > budget:
            break
    ms = 1e3 * sum(times) / len(times)
    v = 1e3 / ms
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "evals/s", "n_gpus": args.gpus, "steps": len(times),
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.rows),
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "note": f"every step is one evaluation over all N rows (no extrapolation); steps are capped by a {budget:.0f} s budget",
            "cpu_baseline": {"value": v, "unit": "evals/s", "cores": threads, "kind": "port",
                             "sample": f"oracle/sgpr.py chunked bound+grad on all {args.rows} rows, {len(times)} evaluation(s) of {ms / 1e3:.1f} s"},
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------------
# secondary legs (GPU + the oracle timed beside each)
# ---------------------------------------------------------------------------------------------------------------------------
def timed(fn, iters, warm):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cpu_time(fn, repeats=2):
    fn()
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def hmc_leg(dev, eng_cls, with_cpu, iters=40, warm=5, chains=4, n_leapfrog=10):
    """HMC / NUTS samples/s on BASELINE configs[1] (co2-shaped N=545, M=100, 4 chains) over theta on the collapsed bound with pymc3's
    priors/transforms (models/bayesian_sgpr_hmc.py:60-78); every leapfrog is ONE batched bound+grad evaluation of all chains."""
    import torch
    import ggp_b200.synthetic as syn
    from ggp_b200.functions import sgpr_vfe_logp_dlogp
    from ggp_b200.hmc import hmc_sample, nuts_sample
    c = syn.config2_co2_shaped()
    X, y, Z = (torch.tensor(c[k], device=dev) for k in ("X", "y", "Z"))
    eng = eng_cls.get(dev)
    D = X.shape[1]
    g = torch.Generator(device=dev).manual_seed(173)
    x0 = torch.zeros(chains, D + 2, dtype=torch.float64, device=dev)
    x0[:, :D] = 0.6931471805599453
    f = lambda xx: sgpr_vfe_logp_dlogp(xx, X, y, Z, engine=eng, group=False)
    out = {"workload": "configs[1]: co2-shaped N=545 D=1 M=100, 4 chains in lock-step on the VFE bound + pymc3 priors", "chains": chains}
    fixed = {"leapfrogs_per_sample": n_leapfrog}
    for mode in ("eager", "cuda_graph"):
        ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

        def progress(it):
            if it == warm - 1:
                ev[0].record()
        g.manual_seed(173)
        res = hmc_sample(f, x0, iters, tune=warm, n_leapfrog=n_leapfrog, step_size=0.02, adapt_mass=False, generator=g,
                         progress=progress, cuda_graph=(mode == "cuda_graph"))
        ev[1].record()
        torch.cuda.synchronize()
        sec = ev[0].elapsed_time(ev[1]) * 1e-3
        fixed[mode] = {"samples_per_s": chains * iters / sec, "bound_grad_evals_per_s": chains * iters * n_leapfrog / sec,
                       "ms_per_batched_leapfrog": 1e3 * sec / (iters * n_leapfrog), "accept_rate": float(res["accept_rate"].mean().item())}
    out["fixed_length_hmc_L10"] = fixed
    # pm.NUTS() with pymc3's defaults (the sampler the reference calls), BASELINE configs[1]'s own length: 4 chains x (500 tune + 1000
    # draws), timed over the draws.  Tree bookkeeping in csrc/nuts.cuh, one CUDA-graph replay per leapfrog (evaluation + bookkeeping).
    def run_nuts(tune, draws, native):
        g.manual_seed(173)
        ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

        def nprogress(it):
            if it == tune - 1:
                ev[0].record()
        xj = x0 + (torch.rand(chains, D + 2, dtype=torch.float64, device=dev, generator=g) * 2.0 - 1.0)
        res = nuts_sample(f, xj, draws, tune=tune, generator=g, progress=nprogress, cuda_graph=True, native=native)
        ev[1].record()
        torch.cuda.synchronize()
        sec = ev[0].elapsed_time(ev[1]) * 1e-3
        lf = float(res["n_leapfrog"].double().mean().item())
        return xj, lf, {"samples_per_s": chains * draws / sec, "tune": tune, "draws": draws, "mean_leapfrogs_per_sample": lf,
                        "batched_evals_per_transition": res["n_evals_sampling"] / draws,   # lock-step: the deepest chain's tree
                        "ms_per_batched_eval": 1e3 * sec / max(res["n_evals_sampling"], 1),
                        "mean_tree_depth": float(res["tree_depth"].double().mean().item()),
                        "accept_stat": float(res["accept_rate"].mean().item()),
                        "diverging_frac": float(res["diverging"].double().mean().item())}
    xj, lf, out["nuts_pymc3_defaults"] = run_nuts(500, 1000, True)
    out["nuts_pymc3_defaults"]["note"] = ("tree bookkeeping on the device (csrc/nuts.cuh, ggp_nuts_*): a leapfrog of the 4 chains = ONE CUDA-graph "
                                          "replay (evaluation + one bookkeeping launch); the host reads one flag per tree doubling")
    _, _, out["nuts_torch_bookkeeping"] = run_nuts(60, 60, False)
    out["nuts_torch_bookkeeping"]["note"] = "same sampler, per-leaf bookkeeping as ~60 torch calls (native=False); evaluation still a graph replay"
    out["samples_per_s"] = out["nuts_pymc3_defaults"]["samples_per_s"]
    if with_cpu:
        from oracle import priors
        Xc, yc, Zc = X.cpu(), y.cpu(), Z.cpu()
        xc = xj[0].cpu()
        t = cpu_time(lambda: priors.sgpr_vfe_logp_dlogp(xc, Xc, yc, Zc), repeats=5)
        out["cpu_baseline"] = {"logp_dlogp_evals_per_s": 1.0 / t, "nuts_samples_per_s_at_the_same_leapfrogs": 1.0 / (t * max(lf, 1.0)),
                               "cores": os.cpu_count(), "kind": "port", "sample": "oracle/priors.py logp+dlogp (autograd), one chain, best of 5"}
    return out


def config5_leg(dev, eng_cls, rank, world, dist, with_cpu, total_chains=64, n_leapfrog=10):
    """BASELINE configs[4]: Bernoulli-probit classification N=2e5, D=16, M=512, 64 HMC chains on the SGPMC log-density
    (models/sgp_hmc.py:63-83: L=10, step 0.01, SimpleStepSizeAdaptation), chains SHARDED over the ranks (no collective in the
    sampling loop); every leapfrog of a rank's chains is one chain-batched launch sequence (per-chain theta and v)."""
    import torch
    import ggp_b200.synthetic as syn
    from ggp_b200.dist import shard_chains
    from ggp_b200.functions import sgpmc_logp_dlogp
    from ggp_b200.hmc import hmc_sample
    c = syn.config5_classification()
    X, y, Z = (torch.tensor(c[k], device=dev) for k in ("X", "y", "Z"))
    N, D = X.shape
    M = Z.shape[0]
    mine = shard_chains(total_chains, rank, world)
    C = len(mine)
    eng = eng_cls.get(dev)
    g = torch.Generator(device=dev).manual_seed(173 + rank)
    x0 = torch.cat([0.1 * torch.randn(C, M, dtype=torch.float64, device=dev, generator=g),
                    torch.full((C, D + 2), 1.0, dtype=torch.float64, device=dev)], dim=1)

    def f(xx):
        lp, gv, gr = sgpmc_logp_dlogp(xx[:, :M], xx[:, M:], X, y, Z, likelihood="bernoulli", engine=eng)
        return lp, torch.cat([gv, gr], dim=1)
    warm, iters = 1, 2
    ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
    eng.profile_read()

    def progress(it):
        if it == warm - 1:
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            eng.profile_read()
            ev[0].record()
    res = hmc_sample(f, x0, iters, tune=warm, n_leapfrog=n_leapfrog, step_size=0.01, adapt_mass=False, adaptation="simple",
                     num_adaptation_steps=10, generator=g, progress=progress)
    ev[1].record()
    torch.cuda.synchronize()
    _, _, launches = eng.profile_read()
    t = torch.tensor([ev[0].elapsed_time(ev[1])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item()) * 1e-3
    evals = iters * n_leapfrog
    # algorithmic work per chain evaluation: a = L^{-1} k (N M^2), backward G_A L^{-1} (2 N M^2), Gbar += G_A^T a (2 N M^2) -> 5 N M^2 FP64 flops
    flops = 5.0 * N * M * M * C * evals
    out = {"workload": "configs[4]: Bernoulli-probit SGPMC N=2e5 D=16 M=512, 64 chains sharded over the ranks, HMC L=10",
           "chains_total": total_chains, "chains_per_rank": C, "leapfrogs_per_sample": n_leapfrog,
           "samples_per_s": total_chains * iters / sec, "logp_dlogp_evals_per_s": total_chains * evals / sec,
           "ms_per_batched_leapfrog": 1e3 * sec / evals, "accept_rate": float(res["accept_rate"].mean().item()),
           "gpu_launches_per_batched_leapfrog": launches / evals,
           "roofline": {"bound": "tensor (FP64 DMMA)", "achieved_tflops_per_gpu": flops / sec / 1e12,
                        "algorithmic_flops_per_chain_eval": 5.0 * N * M * M}}
    if with_cpu and rank == 0:
        from oracle import sgpmc
        torch.set_num_threads(os.cpu_count() or 1)
        ns = 16384
        Xc, yc, Zc = X[:ns].cpu(), y[:ns].cpu(), Z.cpu()
        v, raw = x0[0, :M].cpu(), x0[0, M:].cpu()
        tt = cpu_time(lambda: sgpmc.sgpmc_logp_dlogp_chunked(v, raw, Xc, yc, Zc, likelihood="bernoulli", chunk=8192), repeats=1)
        full = tt * N / ns
        out["cpu_baseline"] = {"logp_dlogp_evals_per_s": 1.0 / full, "samples_per_s": 1.0 / (full * n_leapfrog), "cores": os.cpu_count(),
                               "kind": "port",
                               "sample": f"oracle/sgpmc.py chunked logp+dlogp of one chain on {ns} of {N} rows ({tt:.1f} s), time scaled to N"}
    return out


def svgp_leg(dev, eng_cls, with_cpu):
    """BASELINE configs[2]: UCI-Power-shaped N=9568 (7654 train), D=4, M=500: SVGP minibatch ELBO fwd+bwd steps/s (batch 1024, the
    DataLoader order of experiments/regression.py:107-108) and the full-batch SGPR bound+grad on the same data."""
    import torch
    import ggp_b200.synthetic as syn
    c = syn.config3_power_shaped()
    X, y, Z = (torch.tensor(c[k], device=dev) for k in ("X", "y", "Z"))
    n, D = X.shape
    M = Z.shape[0]
    eng = eng_cls.get(dev)
    th = torch.tensor(syn.theta_init_gpytorch(D), device=dev)
    gm = torch.Generator().manual_seed(0)
    qm = (1e-3 * torch.randn(M, dtype=torch.float64, generator=gm)).to(dev)
    qL = torch.eye(M, dtype=torch.float64, device=dev)
    batches = syn.minibatch_indices(n, 1024, generator=torch.Generator().manual_seed(1))
    xb = [X[b.to(dev)] for b in batches]
    yb = [y[b.to(dev)] for b in batches]
    it = [0]

    def step():
        i = it[0] % len(xb)
        it[0] += 1
        return eng.svgp_eval(xb[i], yb[i], Z, qm, qL, th, num_data=n, need_grad=True)
    eng.profile_read()
    ms = timed(step, 24, 8)
    _, _, launches = eng.profile_read()
    ms_sgpr = timed(lambda: eng.sgpr_eval(X, y, Z, th, jitter_policy="gpytorch"), 10, 3)
    out = {"workload": "configs[2]: Power-shaped N=7654 (train) D=4 M=500; SVGP minibatch 1024 (ELBO value+gradient) and full-batch SGPR bound+grad",
           "svgp_steps_per_s": 1e3 / ms, "svgp_ms_per_step": ms, "gpu_launches_per_svgp_step": launches / 32.0,
           "sgpr_evals_per_s": 1e3 / ms_sgpr, "sgpr_ms_per_eval": ms_sgpr}
    if with_cpu:
        from oracle import svgp as osv, sgpr as osgpr
        Xc, yc, Zc, thc = X.cpu(), y.cpu(), Z.cpu(), th.cpu()
        xb0, yb0, qmc, qLc = xb[0].cpu(), yb[0].cpu(), qm.cpu(), qL.cpu()

        def cpu_svgp():
            ps = [t.clone().requires_grad_(True) for t in (Zc, qmc, qLc, thc)]
            e = osv.svgp_elbo(xb0, yb0, ps[0], ps[1], ps[2], ps[3][:D], ps[3][D], ps[3][D + 1], n)
            torch.autograd.grad(e, ps)
        t1 = cpu_time(cpu_svgp, repeats=3)
        t2 = cpu_time(lambda: osgpr.sgpr_bound_and_grads_autograd(Xc, yc, Zc, thc[:D], thc[D], thc[D + 1]), repeats=2)
        out["cpu_baseline"] = {"svgp_steps_per_s": 1.0 / t1, "sgpr_evals_per_s": 1.0 / t2, "cores": os.cpu_count(), "kind": "port",
                               "sample": "oracle/svgp.py ELBO + autograd on one 1024-row minibatch; oracle/sgpr.py bound + autograd on all 7654 rows"}
    return out


def config1_leg(dev, eng_cls, with_cpu):
    """BASELINE configs[0]: demo_1d_regression N=1000, M=20, D=1 -- the reference's own CPU-runnable case (models/sgpr.py:168-181)."""
    import torch
    import ggp_b200.synthetic as syn
    c = syn.config1_demo_1d()
    X, y, Z = (torch.tensor(c[k], device=dev) for k in ("X", "y", "Z"))
    eng = eng_cls.get(dev)
    th = torch.tensor(syn.theta_init_gpytorch(1), device=dev)
    eng.profile_read()
    ms = timed(lambda: eng.sgpr_eval(X, y, Z, th, jitter_policy="gpytorch"), 50, 10)
    _, _, launches = eng.profile_read()
    out = {"workload": "configs[0]: demo_1d_regression N=1000 D=1 M=20 SGPR bound+grad", "evals_per_s": 1e3 / ms, "ms_per_eval": ms,
           "gpu_launches_per_eval": launches / 60.0}
    if with_cpu:
        from oracle import sgpr as osgpr
        Xc, yc, Zc, thc = X.cpu(), y.cpu(), Z.cpu(), th.cpu()
        t = cpu_time(lambda: osgpr.sgpr_bound_and_grads_autograd(Xc, yc, Zc, thc[:1], thc[1], thc[2]), repeats=5)
        out["cpu_baseline"] = {"evals_per_s": 1.0 / t, "cores": os.cpu_count(), "kind": "port", "sample": "oracle/sgpr.py bound + autograd, best of 5"}
    return out


def variants_leg(dev, eng_cls, syn, flush, precision):
    """SURVEY 8d variants of the headline configuration, 3 timed evaluations each: Z drawn WITH replacement (the reference's rule,
    experiments/regression.py:83 -- duplicate rows engage the jitter ladder; the sliced-integer engine then evaluates on its FP64
    DMMA plan), theta at the gpytorch initial point, and the strictly streaming mode (tile_cache_mib = 0)."""
    import torch
    out = {}
    th_tr = torch.tensor(syn.theta_trained_like(D_IN), device=dev)
    th_in = torch.tensor(syn.theta_init_gpytorch(D_IN), device=dev)
    c = syn.config4_large()
    X, y, Z = (torch.tensor(c[k], device=dev) for k in ("X", "y", "Z"))
    eng = eng_cls.get(dev, precision=precision)

    def run(e, Zv, thv):
        def f():
            flush.zero_()
            return e.sgpr_eval(X, y, Zv, thv, jitter_policy="gpytorch")
        ms = timed(f, 3, 2)
        o = f()
        return {"evals_per_s": 1e3 / ms, "ms_per_step": ms, "jitter": float(o["jitter"][0].item()), "path": o["path"],
                "bound_value": float(o["bound"][0].item())}
    out["theta_gpytorch_init"] = run(eng, Z, th_in)
    cw = syn.config4_large(with_replacement=True)
    Zw = torch.tensor(cw["Z"], device=dev)
    out["Z_with_replacement"] = run(eng, Zw, th_tr)
    out["Z_with_replacement"]["duplicate_rows"] = int(M_IND - len(set(cw["Z_idx"].tolist())))
    eng0 = eng_cls.get(dev, precision=precision, tile_cache_mib=0)
    out["streaming_tile_cache_0"] = run(eng0, Z, th_tr)
    out["streaming_tile_cache_0"]["note"] = "never more than chunk_rows x m of k(X,Z) / A alive; tiles rebuilt in pass 2"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_FULL, help="override N (debug only; the headline number needs the default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hmc", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the secondary workloads (variants, configs[0,2,4])")
    ap.add_argument("--precision", default=os.environ.get("GGP_BENCH_PRECISION", "fp64_i8"), choices=["fp64", "fp64_i8"],
                    help="fp64_i8 (default): FP64-class exact int8 slicing on tcgen05 (csrc/gemm_i8.cuh); fp64: FP64 DMMA tensor path")
    ap.add_argument("--no-dmma-leg", action="store_true", help="skip the secondary measurement of the FP64 DMMA path")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference(args, rank)

    import numpy as np
    import torch
    import torch.distributed as dist
    import ggp_b200
    import ggp_b200.synthetic as syn

    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N = args.rows
    c = syn.config4_large(N=N, D=D_IN, M=M_IND)
    lo, hi = rank * N // world, (rank + 1) * N // world
    Xh = torch.tensor(c["X"][lo:hi]).pin_memory()
    yh = torch.tensor(c["y"][lo:hi]).pin_memory()
    Zh = torch.tensor(c["Z"]).pin_memory()
    th_np = syn.theta_trained_like(D_IN)
    thh = torch.tensor(th_np).pin_memory()
    X, y, Z, th = Xh.to(dev), yh.to(dev), Zh.to(dev), thh.to(dev)
    n_local = hi - lo
    eng = ggp_b200.Engine.get(dev, precision=args.precision)
    group = True if world > 1 else False   # rows are sharded over the ranks: opt in to the two all-reduces per evaluation
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_resident():
        flush.zero_()
        return eng.sgpr_eval(X, y, Z, th, jitter_policy="gpytorch", need_grad=True, group=group)

    for _ in range(args.warmup):
        out = step_resident()
    torch.cuda.synchronize()
    peak = eng.probe_dmma_peak(20000)  # measured FP64 tensor-pipe peak on this GPU, TFLOP/s
    i8_burst = eng.probe_i8_peak(2000)      # tcgen05 kind::i8 issue loop, ~10 ms per launch: burst
    i8_sust = eng.probe_i8_peak(40000)      # ~0.2 s per launch back to back: sustained under the power cap
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_region():
        eng.profile_read()
        eng.profile_enable(True)
        sampler = ClockSampler(local)
        if rank == 0 and not os.environ.get("GGP_BENCH_NO_SAMPLER"):   # (developer switch: diagnostics only)
            sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        o = None
        for _ in range(args.steps):
            o = step_resident()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        c_ms, c_n, nl = eng.profile_read()
        eng.profile_enable(False)
        return o, ms, c_ms, c_n, nl, (sampler.stop() if rank == 0 else None)

    out, ms_total, cat_ms, cat_n, launches, clocks = timed_region()
    # a timed region that saw a hardware / thermal slowdown is rejected and measured once more (sw_power_cap is kept and reported);
    # rank 0 decides for all ranks so that the collective calls stay matched
    redo = torch.zeros(1, dtype=torch.int32, device=dev)
    if rank == 0 and clocks and any(r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown") for r in clocks.get("reasons", [])):
        redo += 1
    if world > 1:
        dist.broadcast(redo, src=0)
    remeasured = bool(int(redo.item()))
    if remeasured:
        first_reasons = clocks.get("reasons", []) if clocks else []
        out, ms_total, cat_ms, cat_n, launches, clocks = timed_region()
        if clocks is not None:
            clocks["remeasured_after"] = first_reasons
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = 1e3 / ms_step
    head_F = float(out["bound"][0].item())
    head_g = out["grad"][0].cpu().numpy()

    # ---- secondary leg: the same step on the FP64 DMMA path (mma.sync m8n8k4 f64), for comparison with the sliced-integer path
    dmma = None
    dmma_F = dmma_g = None
    if args.precision != "fp64" and not args.no_dmma_leg:
        eng2 = ggp_b200.Engine.get(dev, precision="fp64")

        def step_dmma():
            flush.zero_()
            return eng2.sgpr_eval(X, y, Z, th, jitter_policy="gpytorch", need_grad=True, group=group)
        for _ in range(2):
            o2 = step_dmma()
        torch.cuda.synchronize()
        eng2.profile_read(); eng2.profile_enable(True)
        if world > 1:
            dist.barrier()
        nst = min(args.steps, 5)
        e0.record()
        for _ in range(nst):
            o2 = step_dmma()
        e1.record()
        torch.cuda.synchronize()
        c2, n2, _ = eng2.profile_read(); eng2.profile_enable(False)
        t3 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        ms2 = float(t3.item()) / nst
        g2 = (c2["trmm"] + c2["syrk"] + c2["bwd"]) / nst
        dmma_F, dmma_g = float(o2["bound"][0].item()), o2["grad"][0].cpu().numpy()
        # algorithmic 4 N M^2; the DMMA path executes 5 N M^2 (it rebuilds A = L^{-1} Kzx per chunk in pass 2)
        dmma = {"value": 1e3 / ms2, "unit": "evals/s", "ms_per_step": ms2, "kernel": "k_gemm_tma (TMA-fed DMMA.8x8x4)",
                "gemm_tflops_algorithmic": 4.0 * n_local * M_IND * M_IND / (g2 * 1e-3) / 1e12 if g2 > 0 else None,
                "gemm_tflops_executed": 5.0 * n_local * M_IND * M_IND / (g2 * 1e-3) / 1e12 if g2 > 0 else None,
                "frac_of_dmma_peak_executed": (5.0 * n_local * M_IND * M_IND / (g2 * 1e-3) / 1e12 / peak["best"]) if g2 > 0 else None,
                "bound_value": dmma_F,
                "rel_diff_of_bound_vs_headline_path": abs(dmma_F - head_F) / abs(dmma_F),
                "rel_diff_of_grad_vs_headline_path": float(np.abs(dmma_g - head_g).max() / np.abs(dmma_g).max())}
        del eng2

    # ---- end to end through the public API with HOST buffers: H2D of the step's inputs and D2H of its result inside the timed region
    P = D_IN + 2 + M_IND * D_IN
    res_h = torch.empty(1 + P, dtype=torch.float64).pin_memory()

    def step_e2e():
        flush.zero_()
        # pinned HOST tensors straight into the public call: it uploads Z, theta first and X, y on its side stream (H2D of the rows
        # overlaps the Kzz factorisation); every byte is copied again every step
        o = eng.sgpr_eval(Xh, yh, Zh, thh, jitter_policy="gpytorch", need_grad=True, group=group)
        res_h[:1].copy_(o["bound"], non_blocking=True)
        res_h[1:].copy_(o["grad"][0], non_blocking=True)
        torch.cuda.synchronize()
        return res_h

    step_e2e()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    torch.cuda.synchronize()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.item()) / args.steps
    h2d = (n_local * D_IN + n_local + M_IND * D_IN + D_IN + 2) * 8
    d2h = (1 + P) * 8

    # ---- configs[4]: chains sharded over ALL ranks (runs on every rank; collective only in the timing)
    c5 = None
    if not args.no_legs and not args.no_hmc and N == N_FULL:
        try:
            c5 = config5_leg(dev, ggp_b200.Engine, rank, world, dist, with_cpu=(world == 1 and not args.no_cpu_baseline))
        except Exception as ex:   # a secondary leg must never take the headline line down
            c5 = {"error": repr(ex)}

    if rank == 0:
        traffic_i8 = traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r1b_ncu_traffic.json")) as fh:
                traffic = json.load(fh)["traffic_bytes_per_launch_avg"]
            for name in ("r2d_ncu_traffic_i8.json", "r2_ncu_traffic_i8.json", "r1d_ncu_traffic_i8.json"):
                pth = os.path.join(ROOT, "profiles", name)
                if os.path.exists(pth):
                    with open(pth) as fh:
                        traffic_i8 = json.load(fh)["traffic_bytes_per_launch_avg"]
                    break
        except Exception:
            pass
        gemm_ms = (cat_ms["trmm"] + cat_ms["syrk"] + cat_ms["bwd"]) / args.steps
        gemm_launches = (cat_n["trmm"] + cat_n["syrk"] + cat_n["bwd"]) / args.steps
        flops_local = 4.0 * n_local * M_IND * M_IND  # N M^2 (tri) + N M^2 (syrk) + 2 N M^2 (backward)
        achieved = flops_local / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        if args.precision == "fp64_i8":
            # sliced-integer path: 28 int8 digit-pair products per FP64 product (7 radix-256 digits, pairs i + j <= 6).  Peak = the
            # tcgen05 kind::i8 issue loop measured on THIS GPU in THIS run (ggp_probe_i8_peak: operands resident in shared memory, no
            # loads, no epilogue), the sustained figure because the GEMM spans are timed inside a long power-capped step.
            tops = 28.0 * flops_local / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
            pk = i8_sust["uniform_n256"]
            roof = {"bound": "tensor", "kernel": "k_gemm_i8 (TMA-fed tcgen05.mma kind::i8, int32 TMEM accumulators: triangular multiply + SYRK + backward GEMM)",
                    "achieved": tops, "peak": pk, "unit": "TOP/s (int8)", "frac": tops / pk if tops else None,
                    "peak_source": "measured live: tcgen05.mma kind::i8 128x256x32 issue loop on all SMs, sustained (0.2 s launches back to back)",
                    "peak_burst": i8_burst["uniform_n256"], "frac_of_burst_peak": tops / i8_burst["uniform_n256"] if tops else None,
                    "peak_of_the_production_mma_mix_sustained": i8_sust["mix_depth2"],
                    "frac_of_the_production_mma_mix": tops / i8_sust["mix_depth2"] if tops else None,
                    "peak_of_the_production_mma_mix_burst": i8_burst["mix_depth2"],
                    "fp64_equivalent_tflops": achieved,
                    "fp64_equivalent_vs_dmma_peak": (achieved / peak["best"]) if achieved else None, "dmma_peak_tflops": peak["best"],
                    "digit_products_per_fp64_product": 28, "launches_per_step": gemm_launches,
                    "avg_launch_ms": gemm_ms / gemm_launches if gemm_launches else None,
                    "algorithmic_flops_per_step_per_rank": flops_local, "traffic": traffic_i8,
                    "traffic_unit": "bytes per 16384 rows and GEMM role (dram read+write, ncu --set full, avg of the 3 roles; profiles/)",
                    "whole_step_fp64_equivalent_tflops": flops_local / (ms_step * 1e-3) / 1e12}
            dtype = "f64 via 7 radix-256 int8 digits (exact int32 accumulation on tcgen05, 64-bit integer / f64 recombination)"
        else:
            roof = {"bound": "tensor", "kernel": "k_gemm_tma (TMA-fed DMMA.8x8x4 mainloop: triangular multiply x2 + SYRK + backward GEMM)",
                    "achieved": achieved, "peak": peak["best"], "unit": "TFLOP/s", "frac": (achieved / peak["best"]) if achieved else None,
                    "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, ncu --set full, avg of the 3 GEMM roles on a "
                    "16384-row chunk; profiles/r1b_ncu_traffic.json)", "peak_source": "measured live: register-resident mma.sync m8n8k4 f64 loop on all SMs "
                    "(MEASURED_PEAKS.json holds no FP64 figure)", "launches_per_step": gemm_launches,
                    "avg_launch_ms": gemm_ms / gemm_launches if gemm_launches else None,
                    "algorithmic_flops_per_step_per_rank": flops_local,
                    "whole_step_frac": flops_local / (ms_step * 1e-3) / 1e12 / peak["best"]}
            dtype = "f64"
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": workload_config(N),
            "run": {"rows_per_rank": n_local, "precision": args.precision, "path": out.get("path")},
            "e2e": {"value": 1e3 / e2e_ms, "unit": "evals/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "breakdown_ms_per_step": {k: v / args.steps for k, v in cat_ms.items()},
            "breakdown_note": "CUDA-event spans per category on the stream the category runs on; `build` (side stream) and `mm` overlap: the "
                              "explicit inverse that follows the cluster-resident Kzz factorisation queues behind the build's CTAs, so the mm span "
                              "contains most of the build's duration and the spans do not add up to the step (mm work itself: ~1.4 ms)",
            "breakdown_note": "CUDA-event spans per kernel category; the tile build of pass 1 runs on a side stream next to the Kzz "
                              "factorisation, so the build and mm spans overlap (their sum exceeds their wall time)",
            "bound_value": head_F,
        }
        if dmma is not None:
            line["fp64_dmma_path"] = dmma
        # ---- parity at the headline configuration (full N): against the committed long-double evaluation and the float64 oracle
        parity = {"tolerance": 1e-8, "norm": "max-norm relative, per gradient block"}
        gold = os.path.join(ROOT, "tests", "golden", "c4_headline_ld_without_replacement.npz")
        if N == N_FULL and os.path.exists(gold):
            z = np.load(gold)
            ref_g = {"ell": z["d_ell"], "sf2": float(z["d_sf2"]), "s2": float(z["d_s2"]), "Z": z["d_Z"]}
            parity["vs_long_double"] = {"reference": "oracle/hp (x87 long double) on all 1e6 rows, committed: tests/golden/c4_headline_ld_without_replacement.npz",
                                        args.precision: grad_blocks_relerr(head_g, head_F, float(z["F"]), ref_g, M_IND, D_IN)}
            if dmma_g is not None:
                parity["vs_long_double"]["fp64"] = grad_blocks_relerr(dmma_g, dmma_F, float(z["F"]), ref_g, M_IND, D_IN)
        line["hmc_at_headline_config"] = {"leapfrogs_per_sample": 10, "samples_per_s": value / 10.0,
                                          "note": "one HMC sample = L leapfrogs x one bound+grad evaluation (models/sgp_hmc.py:67-69 uses L=10)"}
        with_cpu = world == 1 and not args.no_cpu_baseline
        if with_cpu:
            threads = os.cpu_count() or 1
            secs, Fo, go, _ = cpu_oracle_full_eval(c, th_np, threads)
            line["cpu_baseline"] = {"value": 1.0 / secs, "unit": "evals/s", "cores": threads, "kind": "port",
                                    "sample": f"oracle/sgpr.py chunked bound+grad on all {N} rows, one evaluation ({secs:.1f} s)"}
            gon = {k: (v.numpy() if hasattr(v, "numpy") else float(v)) for k, v in go.items()}
            parity["vs_float64_oracle"] = {"reference": "oracle/sgpr.py (float64, chunked) on the same inputs, this run",
                                           args.precision: grad_blocks_relerr(head_g, head_F, Fo, gon, M_IND, D_IN)}
            if dmma_g is not None:
                parity["vs_float64_oracle"]["fp64"] = grad_blocks_relerr(dmma_g, dmma_F, Fo, gon, M_IND, D_IN)
        line["parity_at_headline"] = parity
        if c5 is not None:
            line["config5_hmc"] = c5
        if world == 1 and not args.no_legs:
            del X, y
            torch.cuda.empty_cache()
            for name, fn in (("variants", lambda: variants_leg(dev, ggp_b200.Engine, syn, flush, args.precision)),
                             ("svgp", lambda: svgp_leg(dev, ggp_b200.Engine, with_cpu)),
                             ("config1", lambda: config1_leg(dev, ggp_b200.Engine, with_cpu))):
                try:
                    line[name] = fn()
                except Exception as ex:
                    line[name] = {"error": repr(ex)}
        if world == 1 and not args.no_hmc:
            try:
                line["hmc"] = hmc_leg(dev, ggp_b200.Engine, with_cpu)
            except Exception as ex:
                line["hmc"] = {"error": repr(ex)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
