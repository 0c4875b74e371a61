"""Extended-precision SGPR bound + gradient (oracle/hp/sgpr_hp.c through ctypes).  TEST INFRASTRUCTURE ONLY.

`bound_grad(X, y, Z, theta, jitter, precision="ld")` evaluates the closed-form algebra of oracle/sgpr.py in x87 long double
("ld", 64-bit mantissa) or IEEE binary128 ("quad", 113-bit, ~60x slower) on float64 inputs and returns float64-rounded results:
the reference point against which the float64 oracle's own conditioning floor and the GPU paths are measured
(tests/test_oracle_hp.py, tests/test_gpu_headline_parity.py, bench.py `parity_at_headline`).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(os.path.dirname(_HERE), "_ref")
_LIBS = {"ld": "libsgpr_hp_ld.so", "quad": "libsgpr_hp_q.so"}
_loaded = {}


def build(verbose=False):
    """gcc the two variants into oracle/_ref/ (no-op when up to date)."""
    src = os.path.join(_HERE, "sgpr_hp.c")
    if all(os.path.exists(os.path.join(_OUT, f)) and os.path.getmtime(os.path.join(_OUT, f)) >= os.path.getmtime(src)
           for f in _LIBS.values()):
        return _OUT
    subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=None if verbose else subprocess.DEVNULL)
    return _OUT


def _lib(precision):
    if precision not in _loaded:
        path = os.path.join(_OUT, _LIBS[precision])
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.sgpr_hp_bound_grad.restype = ctypes.c_int
        lib.sgpr_hp_bound_grad.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_double, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_void_p]
        lib.sgpr_hp_mantissa_bits.restype = ctypes.c_int
        _loaded[precision] = lib
    return _loaded[precision]


def mantissa_bits(precision="ld"):
    return _lib(precision).sgpr_hp_mantissa_bits()


def bound_grad(X, y, Z, theta, jitter, precision="ld", threads=0):
    """Returns (F, dict(ell[d], sf2, s2, Z[m,d])) as float64 numpy; F is NOT divided by N.  theta = [ell[d], sf2, s2]."""
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    Z = np.ascontiguousarray(np.asarray(Z, dtype=np.float64))
    theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))
    n, d = X.shape
    m = Z.shape[0]
    assert Z.shape[1] == d and theta.shape == (d + 2,) and y.shape == (n,)
    out = np.zeros(3 + d + m * d, dtype=np.float64)
    rc = _lib(precision).sgpr_hp_bound_grad(X.ctypes.data, y.ctypes.data, Z.ctypes.data, theta.ctypes.data, float(jitter), n, m, d,
                                            int(threads or (os.cpu_count() or 1)), out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"sgpr_hp_bound_grad: matrix not positive definite (code {rc})")
    return out[0], dict(ell=out[1:1 + d].copy(), sf2=out[1 + d], s2=out[2 + d], Z=out[3 + d:].reshape(m, d).copy())
