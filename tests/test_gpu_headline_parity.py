"""Parity AT the headline configuration (BASELINE configs[3]: N = 1e6, D = 8, M = 1024; SURVEY 8d: both Z rules, two theta points).

Two layers:
  * reduced N, the SAME inducing set / Kzz: both GPU paths (sliced-integer tcgen05 and FP64 DMMA) and the float64 oracle against the
    long-double evaluation of oracle/hp -- a reference whose own rounding error is ~1e-11 at this conditioning;
  * full N: both GPU paths against the float64 oracle (chunked, ~30 s of host time per case).
Tolerance 1e-8 relative (max-norm per gradient block), the contract of BASELINE.json.  The measured errors are written to
gpurun_out/headline_parity.json for profiles/.
"""
import json
import os

import numpy as np
import pytest
import torch

from helpers import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-8
D, M = 8, 1024
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "headline_parity.json")


def _blocks(bound, grad, Fo, go):
    g = torch.as_tensor(grad, dtype=torch.float64).cpu().reshape(-1)
    return dict(bound=abs(float(bound) - float(Fo)) / abs(float(Fo)), ell=relerr(g[:D], go["ell"]), sf2=relerr(g[D], go["sf2"]),
                s2=relerr(g[D + 1], go["s2"]), Z=relerr(g[D + 2:].view(M, D), go["Z"]))


def _record(key, value):
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        rep = json.load(open(REPORT)) if os.path.exists(REPORT) else {}
        rep[key] = value
        json.dump(rep, open(REPORT, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def _theta(name):
    import ggp_b200.synthetic as syn
    return syn.theta_trained_like(D) if name == "trained" else syn.theta_init_gpytorch(D)


@pytest.mark.parametrize("with_replacement", [False, True])
@pytest.mark.parametrize("theta_name", ["trained", "init"])
def test_reduced_n_same_kzz_against_long_double(with_replacement, theta_name):
    import ggp_b200
    import ggp_b200.synthetic as syn
    from oracle import hp, sgpr as osgpr
    n = 20000   # two chunks, ragged
    c = syn.config4_large(with_replacement=with_replacement)
    X, y, Z, th = c["X"][:n], c["y"][:n], c["Z"], _theta(theta_name)
    Xt, yt, Zt, tht = (torch.tensor(a) for a in (X, y, Z, th))
    dev = torch.device("cuda:0")
    res = {}
    jit = None
    for prec in ("fp64_i8", "fp64"):
        out = ggp_b200.Engine.get(dev, precision=prec).sgpr_eval(Xt, yt, Zt, tht, jitter_policy="gpytorch")
        jit = float(out["jitter"][0].item()) if jit is None else jit
        assert float(out["jitter"][0].item()) == jit
        res[prec] = (float(out["bound"][0].item()), out["grad"][0].cpu())
    F64, g64, jit64 = osgpr.sgpr_bound_and_grads_chunked(Xt, yt, Zt, tht[:D], tht[D], tht[D + 1], "gpytorch", "none", chunk=8192)
    assert jit64 == jit, "the GPU jitter ladder and the oracle's settle on the same level"
    Ft, gt = hp.bound_grad(X, y, Z, th, jit, "ld")
    errs = {k: _blocks(v[0], v[1], Ft, gt) for k, v in res.items()}
    errs["oracle_f64"] = _blocks(F64, torch.cat([g64["ell"], g64["sf2"].reshape(1), g64["s2"].reshape(1), g64["Z"].reshape(-1)]), Ft, gt)
    _record(f"reducedN_{n}_Z_{'with' if with_replacement else 'without'}_replacement_theta_{theta_name}",
            dict(jitter=jit, reference="oracle/hp long double", errors=errs))
    for k, e in errs.items():
        assert max(e.values()) < TOL, (k, e)


@pytest.mark.parametrize("with_replacement", [False, True])
def test_full_n_against_float64_oracle(with_replacement):
    import ggp_b200
    import ggp_b200.synthetic as syn
    from oracle import sgpr as osgpr
    torch.set_num_threads(os.cpu_count() or 1)
    c = syn.config4_large(with_replacement=with_replacement)
    Xt, yt, Zt, tht = [torch.tensor(c[k]) for k in ("X", "y", "Z")] + [torch.tensor(_theta("trained"))]
    dev = torch.device("cuda:0")
    Xd, yd = Xt.to(dev), yt.to(dev)
    res = {}
    for prec in ("fp64_i8", "fp64"):
        out = ggp_b200.Engine.get(dev, precision=prec).sgpr_eval(Xd, yd, Zt, tht, jitter_policy="gpytorch")
        res[prec] = (float(out["bound"][0].item()), out["grad"][0].cpu(), float(out["jitter"][0].item()))
    del Xd, yd
    F64, g64, jit64 = osgpr.sgpr_bound_and_grads_chunked(Xt, yt, Zt, tht[:D], tht[D], tht[D + 1], "gpytorch", "none", chunk=65536)
    errs = {k: _blocks(v[0], v[1], F64, g64) for k, v in res.items()}
    errs["i8_vs_dmma"] = dict(bound=abs(res["fp64_i8"][0] - res["fp64"][0]) / abs(res["fp64"][0]), grad=relerr(res["fp64_i8"][1], res["fp64"][1]))
    _record(f"fullN_1000000_Z_{'with' if with_replacement else 'without'}_replacement_theta_trained",
            dict(jitter=jit64, reference="oracle/sgpr.py float64 (chunked)", errors=errs))
    assert res["fp64_i8"][2] == jit64 and res["fp64"][2] == jit64
    for k in ("fp64_i8", "fp64"):
        assert max(errs[k].values()) < TOL, (k, errs[k])
