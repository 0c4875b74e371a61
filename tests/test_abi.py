"""The C-ABI shared library loads without a GPU and exports every symbol include/ggp_b200.h declares."""
import ctypes
import os
import re

import ggp_b200
from ggp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ggp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ggp_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    ggp_b200.build()
    lib = ctypes.CDLL(ggp_b200.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ggp_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "python binding table and header disagree"
    assert _lib.load().ggp_version() >= 100


def test_workspace_query_needs_no_gpu():
    lib = _lib.load()
    out = ctypes.c_size_t()
    cfg = _lib.GgpCfg(0, 0, 0, 0)
    assert lib.ggp_workspace_bytes(ctypes.byref(cfg), 1_000_000, 1024, 8, 1, ctypes.byref(out)) == 0
    assert 100e6 < out.value < 2e9
    assert lib.ggp_workspace_bytes(ctypes.byref(cfg), 10, 0, 8, 1, ctypes.byref(out)) < 0  # bad argument -> negative code


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "generalised-gaussian-processes_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
