#!/usr/bin/env python
"""Headline benchmark: SGPR collapsed bound + gradient evaluations/s at N=1e6, M=1024, D=8, FP64
(BASELINE.json configs[3], row-sharded over --gpus ranks with one all-reduce per pass).

  python bench.py --gpus 1 --steps K --warmup W              this repo's CUDA path
  python bench.py --impl reference ...                       the CPU restatement of the reference path (oracle/),
                                                             all host threads, bounded row sample, rank 0 only
One JSON line on stdout (rank 0).  A "step" is one bound+gradient evaluation over all N rows.
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


def cpu_oracle_rate(sample_rows, threads, repeats=1):
    """Time the oracle's chunked bound+gradient on `sample_rows` rows of the same workload; return (s per FULL eval, s sample)."""
    import numpy as np
    import torch
    from oracle import sgpr as osgpr
    import ggp_b200.synthetic as syn
    torch.set_num_threads(threads)
    c = syn.config4_large(N=sample_rows, D=D_IN, M=M_IND)
    th = torch.tensor(syn.theta_trained_like(D_IN))
    X, y, Z = (torch.tensor(c[k]) for k in ("X", "y", "Z"))
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        osgpr.sgpr_bound_and_grads_chunked(X, y, Z, th[:D_IN], th[D_IN], th[D_IN + 1], jitter_policy="gpytorch", normalize="n",
                                           chunk=65536)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best * (N_FULL / sample_rows), best


def hmc_rate(dev, eng_cls, iters=40, warm=5, chains=4, n_leapfrog=10):
    """HMC samples/s on BASELINE configs[1] (co2-shaped N=545, M=100, 4 chains): lock-step fixed-L HMC over theta on the collapsed
    bound with pymc3's priors/transforms (models/bayesian_sgpr_hmc.py:60-78); every leapfrog is ONE batched bound+grad evaluation of
    all chains.  Timed over `iters` post-warm-up HMC iterations with CUDA events."""
    import torch
    import ggp_b200.synthetic as syn
    from ggp_b200.functions import sgpr_vfe_logp_dlogp
    from ggp_b200.hmc import hmc_sample
    c = syn.config2_co2_shaped()
    X, y, Z = (torch.tensor(c[k], device=dev) for k in ("X", "y", "Z"))
    eng = eng_cls.get(dev)
    D = X.shape[1]
    g = torch.Generator(device=dev).manual_seed(173)
    x0 = torch.zeros(chains, D + 2, dtype=torch.float64, device=dev)
    x0[:, :D] = 0.6931471805599453
    f = lambda xx: sgpr_vfe_logp_dlogp(xx, X, y, Z, engine=eng, group=False)
    out = {"workload": "configs[1]: co2-shaped N=545 D=1 M=100, 4 chains in lock-step, fixed-length HMC (L=10) on the VFE bound + pymc3 priors",
           "chains": chains, "leapfrogs_per_sample": n_leapfrog}
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
        out[mode] = {"samples_per_s": chains * iters / sec, "bound_grad_evals_per_s": chains * iters * n_leapfrog / sec,
                     "ms_per_batched_leapfrog": 1e3 * sec / (iters * n_leapfrog), "accept_rate": float(res["accept_rate"].mean().item())}
    out["samples_per_s"] = out["cuda_graph"]["samples_per_s"]
    out["note"] = "cuda_graph: the L-leapfrog trajectory (L bound+grad evaluations of all chains) is one CUDA graph replay"
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 131072
    per_full = []
    for i in range(args.warmup + args.steps):
        full_s, _ = cpu_oracle_rate(sample, threads)
        if i >= args.warmup:
            per_full.append(full_s)
    ms = 1e3 * sum(per_full) / len(per_full)
    v = 1e3 / ms
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": v, "unit": "evals/s", "cores": threads, "kind": "port",
                             "sample": f"oracle/sgpr.py chunked bound+grad on {sample} of {N_FULL} rows, time scaled x{N_FULL / sample:.3f}"},
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_FULL, help="override N (debug only; the headline number needs the default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hmc", action="store_true")
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
    thh = torch.tensor(syn.theta_trained_like(D_IN)).pin_memory()
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
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_region():
        eng.profile_read()
        eng.profile_enable(True)
        sampler = ClockSampler(local)
        if rank == 0:
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

    # ---- secondary leg: the same step on the FP64 DMMA path (mma.sync m8n8k4 f64), for comparison with the sliced-integer path
    dmma = None
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
        e0.record()
        for _ in range(args.steps):
            o2 = step_dmma()
        e1.record()
        torch.cuda.synchronize()
        c2, n2, _ = eng2.profile_read(); eng2.profile_enable(False)
        t3 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        ms2 = float(t3.item()) / args.steps
        g2 = (c2["trmm"] + c2["syrk"] + c2["bwd"]) / args.steps
        dmma = {"value": 1e3 / ms2, "unit": "evals/s", "ms_per_step": ms2, "kernel": "k_gemm_tma (TMA-fed DMMA.8x8x4)",
                "gemm_tflops": 4.0 * n_local * M_IND * M_IND / (g2 * 1e-3) / 1e12 if g2 > 0 else None,
                "frac_of_dmma_peak": (4.0 * n_local * M_IND * M_IND / (g2 * 1e-3) / 1e12 / peak["best"]) if g2 > 0 else None,
                "bound_value": float(o2["bound"][0].item()),
                "rel_diff_of_bound_vs_headline_path": abs(float(o2["bound"][0].item()) - float(out["bound"][0].item())) / abs(float(o2["bound"][0].item())),
                "rel_diff_of_grad_vs_headline_path": float(((o2["grad"] - out["grad"]).abs().max() / o2["grad"].abs().max()).item())}
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

    if rank == 0:
        traffic = traffic_i8 = None
        try:
            with open(os.path.join(ROOT, "profiles", "r1b_ncu_traffic.json")) as fh:
                traffic = json.load(fh)["traffic_bytes_per_launch_avg"]
            with open(os.path.join(ROOT, "profiles", "r1d_ncu_traffic_i8.json")) as fh:
                traffic_i8 = json.load(fh)["traffic_bytes_per_launch_avg"]
        except Exception:
            pass
        gemm_ms = (cat_ms["trmm"] + cat_ms["syrk"] + cat_ms["bwd"]) / args.steps
        gemm_launches = (cat_n["trmm"] + cat_n["syrk"] + cat_n["bwd"]) / args.steps
        flops_local = 4.0 * n_local * M_IND * M_IND  # N M^2 (tri) + N M^2 (syrk) + 2 N M^2 (backward)
        achieved = flops_local / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        if args.precision == "fp64_i8":
            # sliced-integer path: 28 int8 digit-pair products per FP64 product (7 radix-256 digits, pairs i + j <= 6); int8 dense tensor peak = 2 x the bf16 dense peak (nominal
            # ratio); the bf16 figure is the cuBLAS burst number the driver measured on this pool (MEASURED_PEAKS.json), else the
            # profiling recipe's fallback 1590 TFLOP/s
            bf16 = float(peaks.get("bf16_tflops", 1590.0))
            src = "measured (MEASURED_PEAKS.json bf16_tflops x 2)" if "bf16_tflops" in peaks else "fallback (1590 bf16 TFLOP/s x 2)"
            tops = 28.0 * flops_local / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
            roof = {"bound": "tensor", "kernel": "k_gemm_i8 (TMA-fed tcgen05.mma kind::i8, int32 TMEM accumulators: triangular multiply + SYRK + backward GEMM)",
                    "achieved": tops, "peak": 2.0 * bf16, "unit": "TOP/s (int8)", "frac": tops / (2.0 * bf16) if tops else None,
                    "peak_source": src, "fp64_equivalent_tflops": achieved,
                    "fp64_equivalent_vs_dmma_peak": (achieved / peak["best"]) if achieved else None, "dmma_peak_tflops": peak["best"],
                    "digit_products_per_fp64_product": 28, "launches_per_step": gemm_launches,
                    "avg_launch_ms": gemm_ms / gemm_launches if gemm_launches else None,
                    "algorithmic_flops_per_step_per_rank": flops_local, "traffic": traffic_i8,
                    "traffic_unit": "bytes per 16384 rows and GEMM role (dram read+write, ncu --set full, avg of the 3 roles; profiles/r1d_ncu_traffic_i8.json)",
                    "whole_step_fp64_equivalent_tflops": flops_local / (ms_step * 1e-3) / 1e12}
            dtype = "f64 via 7 radix-256 int8 digits (exact int32 accumulation on tcgen05, 64-bit integer / f64 recombination); parity 1e-8 as the DMMA path"
        else:
            roof = {"bound": "tensor", "kernel": "k_gemm_tma (TMA-fed DMMA.8x8x4 mainloop: triangular multiply + SYRK + backward GEMM)",
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
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "N": N, "M": M_IND, "D": D_IN, "rows_per_rank": n_local, "theta": "trained-like (ell=sqrt(D), sf2=1, s2=0.1)",
                       "jitter_policy": "gpytorch", "precision": args.precision,
                       "l2": "256 MB buffer written between steps (inside the timed region)"},
            "e2e": {"value": 1e3 / e2e_ms, "unit": "evals/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "breakdown_ms_per_step": {k: v / args.steps for k, v in cat_ms.items()},
            "breakdown_note": "CUDA-event spans per kernel category; the tile build of pass 1 runs on a side stream next to the Kzz "
                              "factorisation, so the build and mm spans overlap (their sum exceeds their wall time)",
            "bound_value": float(out["bound"][0].item()),
        }
        if dmma is not None:
            line["fp64_dmma_path"] = dmma
        line["hmc_at_headline_config"] = {"leapfrogs_per_sample": 10, "samples_per_s": value / 10.0,
                                          "note": "one HMC sample = L leapfrogs x one bound+grad evaluation (models/sgp_hmc.py:67-69 uses L=10)"}
        if world == 1 and not args.no_hmc:
            line["hmc"] = hmc_rate(dev, ggp_b200.Engine)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = 196608 if N >= 196608 else N
            full_s, samp_s = cpu_oracle_rate(sample, threads)
            full_s *= N / N_FULL
            line["cpu_baseline"] = {"value": 1.0 / full_s, "unit": "evals/s", "cores": threads, "kind": "port",
                                    "sample": f"oracle/sgpr.py chunked bound+grad on {sample} of {N} rows ({samp_s:.1f} s), time scaled to N"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
