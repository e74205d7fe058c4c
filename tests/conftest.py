import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import async_multigrid_b200 as amg  # noqa: E402
from async_multigrid_b200 import hierarchy as H  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


_CUDA_BUILD_ERROR = None


def pytest_configure(config):
    global _CUDA_BUILD_ERROR
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    from oracle import build as obuild
    amg.build.build_host()
    try:
        amg.build.build_cuda()
    except Exception as e:      # no nvcc on this box: the CPU-only suites (oracle, host set-up, matrix I/O) must still run
        _CUDA_BUILD_ERROR = e
    obuild.build_oracle()
    try:
        obuild.build_ref()
    except Exception:
        pass


def pytest_collection_modifyitems(config, items):
    """without the built CUDA library the GPU-marked tests cannot run (there is no CPU fallback): skip them loudly, and
    skip the tests that only need the library to LOAD (C-ABI symbols, host-only probes) as well"""
    if _CUDA_BUILD_ERROR is None and os.path.exists(amg.build.CUDA_LIB):
        return
    why = pytest.mark.skip(reason="libamg_b200.so could not be built here: %s" % (_CUDA_BUILD_ERROR,))
    for it in items:
        if "gpu" in it.keywords or it.fspath.basename in ("test_abi.py", "test_sellu_host.py", "test_async_program.py"):
            it.add_marker(why)


def hierarchy_from_golden(name):
    """(Hierarchy without transfers, fixture dict) from tests/golden/<name>.npz"""
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    L = int(d["num_levels"])
    A, P = [], []
    for l in range(L):
        s = d["A%d_shape" % l]
        A.append(H.CSR(s[0], s[1], d["A%d_indptr" % l], d["A%d_indices" % l], d["A%d_data" % l]))
        if l < L - 1:
            s = d["Pp%d_shape" % l]
            P.append(H.CSR(s[0], s[1], d["Pp%d_indptr" % l], d["Pp%d_indices" % l], d["Pp%d_data" % l]))
    return H.Hierarchy(A, P), d


@pytest.fixture(scope="session", params=["lap5pt_n32", "lap7pt_n12"])
def golden(request):
    return hierarchy_from_golden(request.param)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# Tolerance of the synchronous history tests, relres[k] = ||r_k|| / ||r_0||:
#   |relres_gpu[k] - relres_ref[k]| <= HIST_TOL (absolute, i.e. relative to ||r_0||)   for every k, and
#   |relres_gpu[k] - relres_ref[k]| <= HIST_REL * relres_ref[k]                        for every k (the tail included).
# BASELINE.json asks for 1e-10; observed on the B200: <= 1e-16 absolute (summation order is the only difference between
# the CPU loops and the kernels).  The relative bound is what keeps the tail honest: at relres ~ 1e-9 an absolute 1e-13
# alone would still admit 1e-4 relative.  A rounding-level perturbation 1e-16 |u| of the iterate moves the residual by
# ~1e-16 cond-ish / relres relative, so the relative bound cannot be pushed to 1e-10 at the tail.
HIST_TOL = 1e-13
HIST_REL = 1e-6


def assert_hist_close(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert len(got) == len(want), (what, len(got), len(want))
    d = np.abs(got - want)
    assert np.max(d) <= HIST_TOL, (what, "absolute", float(np.max(d)))
    rel = d / np.maximum(np.abs(want), 1e-300)
    assert np.max(rel) <= HIST_REL, (what, "relative", float(np.max(rel)), int(np.argmax(rel)))
