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
    # with duplicated inducing rows and trained-like theta cond(Kzz + 1e-8 I) = 4.4e10: no float64 evaluation holds 1e-8 on dF/dZ there
    # (the float64 ORACLE is 1.3e-8 from the long-double reference, the GPU 0.5e-8 .. 1.2e-8 depending on the rounding of the m x m
    # section); that one case is bounded at 1e-7 on dF/dZ, everything else -- and every block of every other case -- at 1e-8
    hard = with_replacement and theta_name == "trained"
    for k, e in errs.items():
        assert max(v for kk, v in e.items() if kk != "Z") < TOL, (k, e)
        assert e["Z"] < (1e-7 if hard else TOL), (k, e)


@pytest.mark.parametrize("with_replacement", [False, True])
def test_full_n_against_long_double_and_float64_oracle(with_replacement):
    """FULL headline size.  References: (a) the committed long-double evaluation of all 1e6 rows (tests/golden/c4_headline_ld_*.npz,
    scripts/make_headline_truth.py: ~11 minutes of host time per case, so it is a fixture), (b) the float64 oracle run here.
    Without replacement (the benchmark's Z rule, cond(Kzz) 7.6e7) everything holds 1e-8.  With replacement the draw contains two
    duplicated rows, the ladder settles on 1e-8 and cond(Kzz + jI) = 4.4e10: no float64 evaluation holds 1e-8 on dF/dZ there --
    the float64 ORACLE is 1.4e-8 from the long-double reference, the GPU 2e-8 .. 4e-8 (bound, ell, sf2, s2 stay below 2e-9) -- so
    that case is bounded at 1e-7 against long double and 2e-7 against the float64 oracle, and the measured values are recorded."""
    import ggp_b200
    import ggp_b200.synthetic as syn
    from oracle import sgpr as osgpr
    torch.set_num_threads(os.cpu_count() or 1)
    c = syn.config4_large(with_replacement=with_replacement)
    Xt, yt, Zt, tht = [torch.tensor(c[k]) for k in ("X", "y", "Z")] + [torch.tensor(_theta("trained"))]
    tag = "with" if with_replacement else "without"
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"c4_headline_ld_{tag}_replacement.npz"))
    assert float(z["x_checksum"]) == float(c["X"].sum()) and np.array_equal(z["z_idx_head"], c["Z_idx"][:16]), "fixture is for these inputs"
    Ft, gt = float(z["F"]), dict(ell=z["d_ell"], sf2=float(z["d_sf2"]), s2=float(z["d_s2"]), Z=z["d_Z"])
    dev = torch.device("cuda:0")
    Xd, yd = Xt.to(dev), yt.to(dev)
    res = {}
    for prec in ("fp64_i8", "fp64"):
        out = ggp_b200.Engine.get(dev, precision=prec).sgpr_eval(Xd, yd, Zt, tht, jitter_policy="gpytorch")
        res[prec] = (float(out["bound"][0].item()), out["grad"][0].cpu(), float(out["jitter"][0].item()), out["path"])
    del Xd, yd
    assert res["fp64_i8"][2] == float(z["jitter"]) and res["fp64"][2] == float(z["jitter"])
    # the sliced-integer engine leaves its tcgen05 plan only when the ladder had to engage
    assert res["fp64_i8"][3] == ("fp64" if with_replacement else "fp64_i8")
    F64, g64, jit64 = osgpr.sgpr_bound_and_grads_chunked(Xt, yt, Zt, tht[:D], tht[D], tht[D + 1], "gpytorch", "none", chunk=65536)
    assert jit64 == float(z["jitter"])
    vs_ld = {k: _blocks(v[0], v[1], Ft, gt) for k, v in res.items()}
    vs_ld["oracle_f64"] = _blocks(F64, torch.cat([g64["ell"], g64["sf2"].reshape(1), g64["s2"].reshape(1), g64["Z"].reshape(-1)]), Ft, gt)
    vs_64 = {k: _blocks(v[0], v[1], F64, g64) for k, v in res.items()}
    _record(f"fullN_1000000_Z_{tag}_replacement_theta_trained",
            dict(jitter=jit64, path_of_the_fp64_i8_engine=res["fp64_i8"][3], vs_long_double=vs_ld, vs_float64_oracle=vs_64,
                 i8_vs_dmma=dict(bound=abs(res["fp64_i8"][0] - res["fp64"][0]) / abs(res["fp64"][0]), grad=relerr(res["fp64_i8"][1], res["fp64"][1]))))
    for k in ("fp64_i8", "fp64"):
        assert max(vs_ld[k].values()) < (1e-7 if with_replacement else TOL), (k, vs_ld[k])
        assert max(v for kk, v in vs_ld[k].items() if kk != "Z") < TOL, (k, vs_ld[k])
        assert max(vs_64[k].values()) < (2e-7 if with_replacement else TOL), (k, vs_64[k])
