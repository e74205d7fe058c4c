"""CPU: the host-side input provider (hierarchy.py + host/amg_host.cpp) against scipy."""
import numpy as np
import pytest
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


# ---- BASELINE.json configs[3]: elasticity stand-in (the reference's MFEM beam, src/DMEM_BuildMatrix.cpp:442-719) ----
def _beam_nodes(ex, ey, ez, h):
    nx, ny = ex + 1, ey + 1
    node = np.arange((ex + 1) * (ey + 1) * (ez + 1))
    return node % nx, (node // nx) % ny, node // (nx * ny), nx


def test_elasticity_beam_operator_properties():
    ex, ey, ez, hh = 16, 3, 2, 0.5
    A, b = H.elasticity_beam(ex, ey, ez, hh)
    ix, iy, iz, nx = _beam_nodes(ex, ey, ez, hh)
    n = A.nrows
    assert n == 3 * ix.size and np.array_equal(A.indices[A.indptr[:-1]], np.arange(n))       # 3 dof / node, diag first
    S = A.to_scipy()
    assert abs(S - S.T).max() < 1e-12 * abs(S).max()
    row_len = np.diff(A.indptr)
    interior = np.repeat((ix >= 2) & (ix < ex) & (iy >= 1) & (iy < ey) & (iz >= 1) & (iz < ez), 3)
    assert interior.any() and np.all(row_len[interior] == 81)                                # 27 nodes x 3 components
    assert np.all(row_len[np.repeat(ix == 0, 3)] == 1)                                       # clamped face
    # rigid-body modes are in the kernel of every row whose stencil does not touch the clamped face
    free = np.repeat(ix >= 2, 3)
    x, y, z = ix * hh, iy * hh, iz * hh
    modes = []
    for c in range(3):
        t = np.zeros(n); t[c::3] = 1.0; modes.append(t)
    r = np.zeros(n); r[0::3] = -y; r[1::3] = x; modes.append(r)
    r = np.zeros(n); r[1::3] = -z; r[2::3] = y; modes.append(r)
    r = np.zeros(n); r[0::3] = z; r[2::3] = -x; modes.append(r)
    for m in modes:
        assert np.max(np.abs((S @ m)[free])) < 1e-10 * abs(S).max()
    # constant-strain patch test inside material 1
    ux = np.zeros(n); ux[0::3] = x
    inside = np.repeat((ix >= 2) & (ix <= ex // 2 - 2) & (iy >= 1) & (iy < ey) & (iz >= 1) & (iz < ez), 3)
    assert inside.any() and np.max(np.abs((S @ ux)[inside])) < 1e-10 * abs(S).max()
    # material 1 is 50x stiffer than material 2 (src/DMEM_BuildMatrix.cpp:540-547)
    d = A.diagonal()
    k1 = 3 * (2 + nx * (1 + (ey + 1) * 1)); k2 = 3 * (ex - 2 + nx * (1 + (ey + 1) * 1))
    assert abs(d[k1] / d[k2] - 50.0) < 1e-9
    # SPD, and the load is the traction integrated over the face x = L
    import scipy.sparse.linalg as sla
    assert sla.eigsh(S.tocsc(), k=1, sigma=0, return_eigenvectors=False)[0] > 0
    assert abs(b.sum() + 1e-2 * (ey * hh) * (ez * hh)) < 1e-14 and np.all(b[0::3] == 0) and np.all(b[1::3] == 0)


def test_systems_amg_keeps_functions_apart_and_converges():
    """num_functions = 3 (src/DMEM_BuildMatrix.cpp:470): interpolation never mixes displacement components; the
    Chebyshev-accelerated BPX cycle (configs[3]) then converges to 1e-9 in the oracle"""
    from oracle import oracle as O
    A, b = H.elasticity_beam(16, 2, 2)
    h = H.amg_setup(A, num_functions=3, theta=0.5)
    assert h.num_levels >= 3
    func = np.arange(A.nrows) % 3
    for l in range(h.num_levels - 1):
        P = h.P_plain[l]
        rows = np.repeat(np.arange(P.nrows), np.diff(P.indptr))
        cfunc = func[h.cpts[l]]
        assert np.array_equal(func[rows], cfunc[P.indices])
        func = cfunc
    h.build_transfers(H.BPX, 0.6)
    p = O.Problem(h, H.BPX, H.JACOBI, 0.6)
    lo, hi = p.eigs_power(300)
    _, hist, _ = p.solve_sync(b, 1e-9, 800, cheby=((hi + lo) / (hi - lo), 2.0 / (hi + lo)))
    assert hist[-1] < 1e-9


def test_tiled_renumbering_is_the_same_hierarchy():
    """hierarchy.reorder_hierarchy: every level renumbered tile by tile; the oracle's solve of the permuted problem is the
    permuted solve (same cycle count, history to rounding)"""
    from oracle import oracle as O
    n = 12
    A = H.laplacian("7pt", n)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    p0 = H.tiled_permutation(n, n, n, 4)
    assert np.array_equal(np.sort(p0), np.arange(n ** 3))
    h2, perms = H.reorder_hierarchy(h, p0)
    assert all(np.all(np.diff(c) > 0) for c in h2.cpts)                                   # coarse points still follow their fine points
    for l in range(h.num_levels):
        assert np.array_equal(h2.A[l].indices[h2.A[l].indptr[:-1]], np.arange(h.n[l]))   # diag first
    h.build_transfers(H.MULTADD, 0.9)
    h2.build_transfers(H.MULTADD, 0.9)
    b2 = np.empty_like(b)
    b2[p0] = b
    u, hist, _ = O.Problem(h, H.MULTADD, H.JACOBI, 0.9).solve_sync(b, 1e-9, 100)
    u2, hist2, _ = O.Problem(h2, H.MULTADD, H.JACOBI, 0.9).solve_sync(b2, 1e-9, 100)
    assert len(hist) == len(hist2) and np.max(np.abs(hist - hist2)) <= 1e-13
    assert np.max(np.abs(u2[p0] - u)) <= 1e-12 * np.max(np.abs(u))


def test_dmem_driver_defaults_rhs_and_jacobi_weight():
    """host side of the DMEM driver's defaults (SURVEY.md 5.9j,k): per-rank srand(0) right-hand side and the Jacobi weight
    1 / lambda_max(D^-1 A) from 20 CG steps (hypre's estimator restated from its published algorithm -- parity unpinned)"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    b = H.rand_rhs_dmem([0, 5, 9, 9, 12])
    one = H.rand_rhs(5, -0.5, 0.5)
    assert np.array_equal(b[:5], one) and np.array_equal(b[5:9], one[:4]) and np.array_equal(b[9:], one[:3])
    assert np.all(np.abs(b) <= 0.5)
    for prob, n in (("7pt", 10), ("5pt", 24), ("27pt", 7)):
        A = H.laplacian(prob, n)
        hi, lo = H.max_eig_estimate_cg(A, 20)
        d = A.diagonal()
        B = sp.diags(1 / np.sqrt(d)) @ A.to_scipy().tocsr() @ sp.diags(1 / np.sqrt(d))
        true = sla.eigsh(B, k=1, which="LA", return_eigenvectors=False)[0]
        assert lo > 0 and hi <= true * (1 + 1e-12) and hi >= 0.97 * true          # Ritz values lie inside the spectrum
        w = H.dmem_default_smooth_weight(A)
        assert abs(w - 1.0 / hi) < 1e-15 and 0.4 < w < 1.0


# ---- SmoothTransfer against the reference's own object code ---------------------------------------------
def _same_matrix(mine, ip, ix, va, tol):
    assert np.array_equal(mine.indptr, ip) and np.array_equal(mine.indices, ix)        # same pattern AND the same row layout
    assert np.max(np.abs(mine.data - va)) <= tol * max(1.0, np.max(np.abs(va)))


@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
def test_smooth_transfer_matches_reference_fixture(name):
    """amgh_smooth_transfer (the host input provider's Pbar = G P, Rbar = P^T GT) against SmoothTransfer of the reference
    (tests/golden/smooth_transfer.npz: src/SMEM_Setup.cpp compiled unmodified, Eigen replaced by a stand-in): identical
    pattern and row layout (columns descending, diagonal swapped to the front), values to 1e-15"""
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "smooth_transfer.npz")))
    h, d = hierarchy_from_golden(name)
    for tag, kind in (("j", H.JACOBI), ("l1", H.L1_JACOBI)):
        h.build_transfers(H.MULTADD, 0.9, smooth_interp_type=kind)
        for l in range(h.num_levels - 1):
            for mn, mine in (("P", h.P[l]), ("R", h.R[l])):
                k = "%s_%s_%s%d_" % (name, tag, mn, l)
                kp = "%s_%s%d_" % (name, mn, l)
                assert list(g[kp + "shape"]) == [mine.nrows, mine.ncols]
                assert np.array_equal(mine.indptr, g[kp + "indptr"]) and np.array_equal(mine.indices, g[kp + "indices"])
                if k + "data" in g:
                    _same_matrix(mine, g[kp + "indptr"], g[kp + "indices"], g[k + "data"], 1e-15)
                else:                   # big levels: pattern + row / column sums of the values (fixture size)
                    S = mine.to_scipy()
                    assert np.max(np.abs(np.asarray(S.sum(axis=1)).ravel() - g[k + "rowsum"])) <= 1e-14
                    assert np.max(np.abs(np.asarray(S.sum(axis=0)).ravel() - g[k + "colsum"])) <= 1e-14


def test_smooth_transfer_matches_live_reference():
    from oracle import oracle as O
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    for A in (H.laplacian("27pt", 6), H.difconv(8, a=(40.0, -20.0, 10.0), atype=3)):     # symmetric and nonsymmetric operators
        h = H.amg_setup(A)
        for w, kind, pre, post in ((0.8, H.JACOBI, 1, 1), (0.8, H.L1_JACOBI, 1, 1), (0.8, H.JACOBI, 1, 0), (0.8, H.JACOBI, 0, 1)):
            h.build_transfers(H.MULTADD, w, smooth_interp_type=kind, num_pre=pre, num_post=post)
            for l in range(h.num_levels - 1):
                Pb, Rb = O.ref_smooth_transfer(h.A[l], h.P_plain[l], w, kind, pre, post)
                assert (Pb is None) == (post == 0) and (Rb is None) == (pre == 0)
                if Pb is not None:
                    _same_matrix(h.P[l], Pb.indptr, Pb.indices, Pb.data, 1e-15)
                if Rb is not None:
                    _same_matrix(h.R[l], Rb.indptr, Rb.indices, Rb.data, 1e-15)


def test_work_model_and_thread_partition_match_live_reference():
    """ComputeWork, PartitionLevels (BALANCED_THREADS) and PartitionGrids of the reference's object code (src/SMEM_Setup.cpp:590-1170)
    against hierarchy.compute_work / balanced_threads / nnz_balanced_bounds: the work model that sizes the level groups (threads on
    the CPU, CTAs in the persistent kernel), the threads per level and every thread's row range (= the hybrid smoother's blocks)"""
    from oracle import oracle as O
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.laplacian("7pt", 12)
    h = H.amg_setup(A)
    L = h.num_levels
    for solver, base, pre, post in ((H.MULTADD, H.MULTADD, 1, 1), (H.MULTADD, H.MULTADD, 1, 0), (H.AFACX, H.AFACX, 1, 1),
                                    (H.ASYNC_MULTADD, H.MULTADD, 1, 1), (H.ASYNC_AFACX, H.AFACX, 1, 1)):
        h.build_transfers(base, 0.9, num_pre=pre, num_post=post)
        for nt, sweeps in ((L, 1), (8, 1), (16, 2), (37, 1)):
            r = O.ref_work_partition(h, solver, nt, pre, post, sweeps, sweeps)
            work, frac = H.compute_work(h, solver, pre, post, sweeps, sweeps)
            tpl = H.balanced_threads(frac, nt)
            assert list(r["level_work"]) == list(work)
            assert np.array_equal(r["frac"], np.asarray(frac))
            assert list(r["threads_per_level"]) == list(tpl)
            if min(tpl) == 0:
                continue            # the reference hangs in PartitionGrids when a level has no thread (driver comment); nothing to compare
            t = 0
            for k in range(L):
                for l in range(L):
                    b = H.nnz_balanced_bounds(h.A[l].indptr, tpl[k])
                    assert list(r["A_ns"][l][t:t + tpl[k]]) == list(b[:-1]) and list(r["A_ne"][l][t:t + tpl[k]]) == list(b[1:])
                t += tpl[k]


def test_stencil_coefficients_match_live_reference():
    """`-problem 7pt | 27pt | difconv`: the coefficients src/BuildHypreMatrix.cpp:100-289 computes (the reference's object code; hypre's
    generators, un-vendored, are replaced by recorders) against an interior row of the host generators' matrices"""
    from oracle import oracle as O
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    n = 8
    i = (n // 2) * (1 + n + n * n)

    def interior(A):
        row = A.to_scipy().tocsr().getrow(i).toarray().ravel()
        return np.array([row[i], row[i - 1], row[i - n], row[i - n * n], row[i + 1], row[i + n], row[i + n * n]])
    m = interior(H.laplacian("7pt", n))
    r = O.ref_stencil_values(1, n, (1.0, 1.0, 1.0))                    # SMEM_Main.cpp:199-204: cx = cy = cz = 1, a = 0
    assert list(r) == [m[0], m[1], m[2], m[3]] and list(m[1:4]) == list(m[4:])
    row = H.laplacian("27pt", n).to_scipy().tocsr().getrow(i).toarray().ravel()
    r = O.ref_stencil_values(2, n)
    assert r[0] == row[i] and np.all(row[np.nonzero(row)[0][np.nonzero(row)[0] != i]] == r[1]) and np.count_nonzero(row) == 27
    for c, a, atype in (((1.0, 1.0, 1.0), (1.0, 1.0, 1.0), -1), ((1.0, 1.0, 1.0), (40.0, -20.0, 10.0), 3), ((1.0, 2.0, 3.0), (10.0, 10.0, 10.0), 0),
                        ((1.0, 1.0, 1.0), (5.0, -7.0, 9.0), 1), ((1.0, 1.0, 1.0), (5.0, -7.0, 9.0), 2), ((0.5, 1.0, 2.0), (-3.0, 4.0, -5.0), 3)):
        r = O.ref_stencil_values(7, n, c, a, atype)
        m = interior(H.difconv(n, c=c, a=a, atype=atype))
        assert np.max(np.abs(r - m)) <= 1e-13 * np.max(np.abs(r)), (atype, r, m)


def test_processor_grid_matches_live_reference():
    """the reference's processor-grid search (src/BuildHypreMatrix.cpp:36-76, object code) against its restatement: y-slabs for a prime
    number of ranks, z-slabs otherwise"""
    from oracle import oracle as O
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    for dims in ((512, 512, 512), (16, 16, 4), (8, 8, 8)):
        for P in (1, 2, 3, 4, 5, 6, 8, 12, 16, 32):
            want = H.reference_processor_grid(P, *dims)
            if want[0] * want[1] * want[2] != P or want[0] > dims[0] or want[1] > dims[1] or want[2] > dims[2]:
                continue        # the reference prints "Invalid number of processors" and exit(1)s for such a count: do not call it
            g = np.zeros(3, dtype=np.int32)
            O.ref_lib().ref_processor_grid(P, dims[0], dims[1], dims[2], O.iptr(g))
            assert tuple(int(v) for v in g) == want, (P, dims)
    assert H.reference_processor_grid(2, 512, 512, 512) == (1, 2, 1) and H.reference_processor_grid(8, 512, 512, 512) == (1, 1, 8)
