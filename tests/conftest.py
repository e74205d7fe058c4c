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


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    from oracle import build as obuild
    amg.build.build_host()
    amg.build.build_cuda()
    obuild.build_oracle()
    try:
        obuild.build_ref()
    except Exception:
        pass


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


# tolerance of the synchronous history test: |relres_gpu[k] - relres_ref[k]| <= HIST_TOL where
# relres = ||r_k|| / ||r_0||  ("within 1e-10 relative", BASELINE.json north_star; relative to r_0:
# two CPU runs with different summation order already differ by 1e-8 relative to ||r_k|| itself
# once ||r_k|| ~ 1e-9 ||r_0||)
HIST_TOL = 1e-10
