"""ctypes binding of libggp_b200.so (the C ABI in include/ggp_b200.h).

There is NO CPU fallback: if the CUDA library is missing or a call fails this module raises.
"""
import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GGP_B200_LIB") or os.path.join(_HERE, "libggp_b200.so")  # env override: developer A/B builds only
CSRC = os.path.join(_HERE, "csrc")

c_double_p = ctypes.c_void_p  # device pointers are passed as integers
c_int_p = ctypes.c_void_p


class GgpCfg(ctypes.Structure):
    _fields_ = [("kernel", ctypes.c_int32), ("precision", ctypes.c_int32),
                ("chunk_rows", ctypes.c_int32), ("tile_cache_mib", ctypes.c_int32), ("kernel_param", ctypes.c_double)]


# ggp_nuts_state (include/ggp_b200.h): field order is the header's
NUTS_DOUBLE_CP = ("x", "g", "inv_mass", "x_eval", "g_eval", "xl", "pl", "gl", "xr", "pr", "gr", "x_prop", "g_prop", "p_sum",
                  "xe", "pe", "ge", "p_half", "s_p_sum", "s_x", "s_g")                              # [C,P] float64
NUTS_DOUBLE_C = ("lp", "eps", "lp_eval", "e0", "lp_prop", "log_w", "sum_acc", "n_leaf", "e", "s_log_w", "s_lp", "acc_prob")   # [C]
NUTS_INT_C = ("depth", "diverged", "active", "right", "s_turn", "s_div", "building", "leaf")       # [C] int32
_NUTS_PTRS = ("x", "lp", "g", "eps", "inv_mass", "x_eval", "lp_eval", "g_eval",
              "e0", "xl", "pl", "gl", "xr", "pr", "gr", "x_prop", "lp_prop", "g_prop", "log_w", "p_sum", "sum_acc", "n_leaf",
              "depth", "diverged", "active",
              "e", "xe", "pe", "ge", "p_half", "s_log_w", "s_p_sum", "s_x", "s_lp", "s_g", "p_ck", "ps_ck",
              "right", "s_turn", "s_div", "building", "leaf", "any_active", "u",
              "acc_prob", "samples", "lps", "depths", "nleaps", "divs")


class GgpNutsState(ctypes.Structure):
    _fields_ = ([("C", ctypes.c_int32), ("P", ctypes.c_int32), ("K", ctypes.c_int32), ("pad_", ctypes.c_int32),
                 ("max_energy_error", ctypes.c_double)] + [(n, ctypes.c_void_p) for n in _NUTS_PTRS])


KERNELS = {"rbf": 0, "matern32": 1, "matern52": 2, "rq": 3}
FACTOR_KINDS = {"rbf": 0, "matern32": 1, "matern52": 2, "rq": 3, "periodic": 4}    # factors of a composite kernel (ggp_kprog.kind)
KERNEL_COMPOSITE = 5
KPROG_MAX_TERMS, KPROG_MAX_FACTORS, KPROG_MAX_PARAMS = 6, 3, 64


class GgpKprog(ctypes.Structure):
    _fields_ = [("nterms", ctypes.c_int32), ("nfactors", ctypes.c_int32 * KPROG_MAX_TERMS),
                ("kind", (ctypes.c_int32 * KPROG_MAX_FACTORS) * KPROG_MAX_TERMS)]


def make_kprog(prog):
    """prog: tuple of terms, each a tuple of factor names ("rbf", "matern32", "matern52", "rq", "periodic"); the parameter row is, in
    program order, a_t then for each factor ell[d] then (rq: alpha | periodic: period[d])  (include/ggp_b200.h)."""
    prog = tuple(tuple(t) for t in prog)
    if not 1 <= len(prog) <= KPROG_MAX_TERMS or any(not 1 <= len(t) <= KPROG_MAX_FACTORS for t in prog):
        raise ValueError(f"composite kernel: 1..{KPROG_MAX_TERMS} terms of 1..{KPROG_MAX_FACTORS} factors")
    kp = GgpKprog()
    kp.nterms = len(prog)
    for t, term in enumerate(prog):
        kp.nfactors[t] = len(term)
        for f, name in enumerate(term):
            kp.kind[t][f] = FACTOR_KINDS[name]
    return kp


def kprog_layout(prog, d):
    """(P, amplitude indices) of the parameter row of `prog` in d input dimensions."""
    npar = {"rbf": d, "matern32": d, "matern52": d, "rq": d + 1, "periodic": 2 * d}
    amp, p = [], 0
    for term in prog:
        amp.append(p)
        p += 1 + sum(npar[f] for f in term)
    return p, amp
PRECISIONS = {"fp64": 0, "tf32x3": 1, "fp64_i8": 2}
LIKELIHOODS = {"gaussian": 0, "bernoulli": 1}

# every symbol include/ggp_b200.h declares: name -> (restype, argtypes)
_I, _I64, _D, _P, _SZP = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)
_CFG = ctypes.POINTER(GgpCfg)
SYMBOLS = {
    "ggp_version": (_I, []),
    "ggp_last_error": (ctypes.c_char_p, []),
    "ggp_create": (_I, [ctypes.POINTER(ctypes.c_void_p), _I]),
    "ggp_destroy": (_I, [_P]),
    "ggp_workspace_bytes": (_I, [_CFG, _I64, _I, _I, _I, _SZP]),
    "ggp_reserve": (_I, [_P, _CFG, _I64, _I, _I, _I]),
    "ggp_sgpr_factor": (_I, [_P, _CFG, _P, _P, _P, _P, _I, _I, _I, _P]),
    "ggp_sgpr_prefetch_tiles": (_I, [_P, _CFG, _P, _P, _I64, _P, _P, _I, _I, _I]),
    "ggp_sgpr_prefetch_tiles_part": (_I, [_P, _CFG, _P, _P, _I64, _I64, _I64, _P, _P, _I, _I, _I]),
    "ggp_sgpr_pass1": (_I, [_P, _CFG, _P, _P, _P, _I64, _P, _P, _I, _I, _I, _P]),
    "ggp_sgpr_predict_pass1": (_I, [_P, _CFG, _P, _P, _P, _I64, _P, _P, _I, _I, _I, _P]),
    "ggp_sgpr_finish": (_I, [_P, _CFG, _P, _P, _P, _I, _I, _I, _P, _I, _P, _P, _P]),
    "ggp_sgpr_join": (_I, [_P, _P]),
    "ggp_sgpr_expect_prefetch": (_I, [_P, _I]),
    "ggp_sgpr_pass2": (_I, [_P, _CFG, _P, _P, _P, _I64, _P, _P, _I, _I, _I, _P]),
    "ggp_sgpr_predict": (_I, [_P, _CFG, _P, _P, _I64, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "ggp_svgp_elbo": (_I, [_P, _CFG, _P, _P, _P, _I64, _P, _P, _I, _P, _P, _P, _I, _I, _I, _D, _D, _D, _I, _I, _P, _P, _P]),
    "ggp_svgp_predict": (_I, [_P, _CFG, _P, _P, _I64, _P, _P, _I, _P, _P, _P, _I, _I, _I, _D, _I, _P, _P, _P]),
    "ggp_chol_batched": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "ggp_gemm_nt": (_I, [_P, _P, _P, _I64, _P, _I64, _P, _I64, _I, _I, _I, _D, _D]),
    "ggp_gemm_nt_ex": (_I, [_P, _P, _P, _I64, _P, _I64, _P, _I64, _I, _I, _I, _D, _D, _I, _I, _I, _I64]),
    "ggp_gemm_nt_i8": (_I, [_P, _P, _P, _I64, _P, _I64, _P, _I64, _I, _I, _I]),
    "ggp_kernel_matrix": (_I, [_P, _CFG, _P, _P, _I64, _P, _I64, _P, _I, _P]),
    "ggp_profile_enable": (_I, [_P, _I]),
    "ggp_profile_read": (_I, [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "ggp_probe_dmma_peak": (_I, [_P, _P, _I, ctypes.POINTER(ctypes.c_double)]),
    "ggp_probe_i8_peak": (_I, [_P, _P, _I, ctypes.POINTER(ctypes.c_double)]),
    "ggp_kprog_nparams": (_I, [ctypes.POINTER(GgpKprog), _I]),
    "ggp_set_kernel_program": (_I, [_P, ctypes.POINTER(GgpKprog), _I]),
    "ggp_set_kernel_params": (_I, [_P, _P, _P, _P]),
    "ggp_vfe_theta": (_I, [_P, _P, _I, _I, _P]),
    "ggp_vfe_logp": (_I, [_P, _P, _P, _P, _I64, _P, _P, _I, _I, _I, _P, _P]),
    "ggp_nuts_state_size": (_I, []),
    "ggp_nuts_begin": (_I, [_P, ctypes.POINTER(GgpNutsState), _P]),
    "ggp_nuts_subtree_begin": (_I, [_P, ctypes.POINTER(GgpNutsState)]),
    "ggp_nuts_leaf": (_I, [_P, ctypes.POINTER(GgpNutsState)]),
    "ggp_nuts_subtree_end": (_I, [_P, ctypes.POINTER(GgpNutsState)]),
    "ggp_nuts_end": (_I, [_P, ctypes.POINTER(GgpNutsState), _I]),
}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]


def build(verbose=False):
    """Compile csrc/ggp_api.cu for sm_100a into libggp_b200.so (in-tree, travels with gpurun)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "ggp_api.cu")]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built (python -c 'import __graft_entry__ as g; "
            "g.build()').  There is no CPU fallback for the sparse-GP hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.ggp_nuts_state_size() != ctypes.sizeof(GgpNutsState):
        raise RuntimeError("ggp_nuts_state: the ctypes mirror in _lib.py and include/ggp_b200.h disagree")
    _lib = lib
    return lib


class GgpError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = load().ggp_last_error()
        raise GgpError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
