"""CPU: the C-ABI shared library loads and exports every symbol include/amg_b200.h declares; no
compute call succeeds without a GPU (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

import async_multigrid_b200 as amg
from conftest import ROOT, has_gpu


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "amg_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(amgb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(amg.build.CUDA_LIB)
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert set(amg.solver.ABI_SYMBOLS) <= set(syms)


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    lib = amg.solver.load_library()
    ctx = ctypes.c_void_p()
    assert lib.amgb_create(ctypes.byref(ctx), 0) != 0
    from async_multigrid_b200 import hierarchy as H
    h = H.amg_setup(H.laplacian("5pt", 8))
    h.build_transfers(H.MULTADD, 1.0)
    with pytest.raises(amg.AmgError):
        amg.Solver(h)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "async-multigrid_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f
