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


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/amg_b200.h compiles as C99 and a C program links against libamg_b200.so -- the binding a
    maintainer of the (C-style C++) reference would write needs nothing else."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "amg_b200.h"\n#include <stdio.h>\n'
                   'int main(void){ amgb_ctx *c = 0; amgb_options o; amgb_default_options(&o);\n'
                   '  int rc = amgb_create(&c, 0);\n'
                   '  printf("%d %d %d\\n", rc, o.solver, o.num_pre_smooth_sweeps);\n'
                   '  if (rc == AMGB_OK) amgb_destroy(c);\n'
                   '  return 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(amg.build.CUDA_LIB)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lamg_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rc, solver, pre = [int(x) for x in out.stdout.split()]
    assert solver == 2 and pre == 1                       # MULTADD, the reference's defaults
    assert (rc == 0) == has_gpu()                         # no device -> an error code, never a CPU fallback
