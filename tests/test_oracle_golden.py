"""CPU: the oracle restatement (oracle/amg_oracle.c) against outputs of the reference's own object
code (committed fixtures, and live oracle/_ref when it is present)."""
import ctypes as C

import numpy as np
import pytest

from conftest import HIST_TOL
from async_multigrid_b200 import hierarchy as H
from oracle import oracle as O


def _close_hist(a, b):
    assert len(a) == len(b), (len(a), len(b))
    assert np.max(np.abs(np.asarray(a) - np.asarray(b))) <= HIST_TOL


def test_multadd_symmetric_jacobi_history(golden):
    h, d = golden
    w = float(d["smooth_weight"])
    h.build_transfers(H.MULTADD, w)
    u, hist, _ = O.Problem(h, H.MULTADD, H.JACOBI, w).solve_sync(d["b"], 1e-9, 100)
    _close_hist(hist, d["multadd_symj_hist"])
    assert np.max(np.abs(u - d["multadd_symj_u"])) <= 1e-12 * np.max(np.abs(u))


def test_multadd_plain_jacobi_history(golden):
    h, d = golden
    w = float(d["smooth_weight"])
    h.build_transfers(H.MULTADD, w, num_pre=1, num_post=0)
    _, hist, _ = O.Problem(h, H.MULTADD, H.JACOBI, w, num_pre=1, num_post=0).solve_sync(d["b"], 1e-9, 60)
    _close_hist(hist, d["multadd_j_hist"])


def test_multadd_symmetric_l1_history(golden):
    h, d = golden
    w = float(d["smooth_weight"])
    h.build_transfers(H.MULTADD, w, smooth_interp_type=H.L1_JACOBI)
    _, hist, _ = O.Problem(h, H.MULTADD, H.L1_JACOBI, w).solve_sync(d["b"], 1e-9, 100)
    _close_hist(hist, d["multadd_syml1_hist"])


def test_afacx_history(golden):
    h, d = golden
    h.build_transfers(H.AFACX, 0.6)
    _, hist, _ = O.Problem(h, H.AFACX, H.JACOBI, 0.6).solve_sync(d["b"], 1e-9, 40)
    _close_hist(hist, d["afacx_j_hist"])


def test_multiplicative_vcycle_history(golden):
    """MULT: SMEM_Sync_Parfor_Vcycle run by the reference's own SMEM_Solve (fixture), V(1,1), weight 0.8"""
    h, d = golden
    h.build_transfers(H.MULT, 0.8)
    _, hist, _ = O.Problem(h, H.MULT, H.JACOBI, 0.8).solve_sync(d["b"], 1e-9, 100)
    _close_hist(hist, d["mult_j_hist"])


def test_bpx_history(golden):
    h, d = golden
    h.build_transfers(H.AFACX, 0.6)
    _, hist, _ = O.Problem(h, H.BPX, H.JACOBI, 0.6).solve_sync(d["b"], 1e-30, 10)
    ref = d["bpx_j_hist"]
    assert len(hist) == len(ref)
    # BPX is not a convergent stationary iteration: compare relative to the growing residual
    assert np.max(np.abs(hist - ref) / ref) <= 1e-10


def test_kernels_against_reference_objects(golden):
    h, d = golden
    w = float(d["smooth_weight"])
    y = O.spgemv(h.A[0], d["x"], None, 1.0, 0.0)
    assert np.max(np.abs(y - d["matvec_A0_x"])) == 0.0      # same loop order -> bit-identical
    u = O.smooth("symmetric_jacobi", h.A[0], d["b"], w=w, sweeps=1, zero_flag=1)
    assert np.max(np.abs(u - d["seq_symj_b"])) <= 1e-15 * np.max(np.abs(u))
    u = O.smooth("jacobi", h.A[0], d["b"], w=w, sweeps=3, zero_flag=1)
    assert np.max(np.abs(u - d["seq_j3_b"])) <= 1e-15 * np.max(np.abs(u))


@pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_objects_match_oracle():
    A = H.laplacian("5pt", 24)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    h.build_transfers(H.MULTADD, 0.8)
    rs = O.RefSolver(h, H.MULTADD, H.JACOBI, b, 0.8, one_thread_per_level=True)
    if rs.num_threads > O.ref_lib().ref_max_threads():
        pytest.skip("more levels than cores: the reference's spin barriers would oversubscribe")
    out = rs.solve_sync_det(50, 1e-9)
    rs.close()
    _, hist, _ = O.Problem(h, H.MULTADD, H.JACOBI, 0.8).solve_sync(b, 1e-9, 50)
    _close_hist(hist, out["hist"])


def test_hybrid_jgs_block_semantics():
    """block GS == plain Gauss-Seidel when one block covers everything; == Jacobi (w=1) for 1-row blocks"""
    A = H.laplacian("7pt", 6)
    f = H.rand_rhs(A.nrows)
    n = A.nrows
    gs = O.smooth("hybrid_jgs", A, f, sweeps=1, zero_flag=1, blocks=np.asarray([0, n], dtype=np.int32))
    S = A.to_scipy()
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    ref = sla.spsolve_triangular(sp.tril(S).tocsr(), f)
    assert np.max(np.abs(gs - ref)) <= 1e-13
    jac = O.smooth("hybrid_jgs", A, f, sweeps=1, zero_flag=1, blocks=np.arange(n + 1, dtype=np.int32))
    assert np.max(np.abs(jac - f / A.diagonal())) <= 1e-15
    # second sweep, 4-row blocks, against a direct numpy transcription
    blocks = H.uniform_blocks(n, 4)
    got = O.smooth("hybrid_jgs", A, f, sweeps=2, zero_flag=1, blocks=blocks)
    D = S.toarray()
    u = np.zeros(n)
    for k in range(2):
        up = u.copy()
        for b in range(len(blocks) - 1):
            ns, ne = blocks[b], blocks[b + 1]
            if k == 0:
                u[ns:ne] = 0
            for i in range(ns, ne):
                res = f[i]
                for j in np.nonzero(D[i])[0]:
                    if ns <= j < ne:
                        res -= D[i, j] * u[j]
                    elif k > 0:
                        res -= D[i, j] * up[j]
                u[i] = res / D[i, i] if k == 0 else u[i] + res / D[i, i]
    assert np.max(np.abs(got - u)) <= 1e-13


# ---- implicit extended-system BPX (`-solver iebpx`, src/SMEM_ExtendedSystem.cpp) ---------------------------------------
@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
@pytest.mark.parametrize("sm,tag", [(H.JACOBI, "j"), (H.L1_JACOBI, "l1")])
def test_iebpx_matches_reference_fixture(name, sm, tag):
    """oracle restatement vs the reference's own object code (tests/golden/iebpx.npz, made by make_golden.py from
    oracle/_ref): same iteration count at tol 1e-9, same norms, same solution"""
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "iebpx.npz")))
    h, d = hierarchy_from_golden(name)
    h.build_transfers(H.BPX, 0.8)
    pb = O.Problem(h, H.BPX, sm, 0.8)
    for nc in (2, 7, 300):
        k = "%s_%s_nc%d_" % (name, tag, nc)
        mu, delta = g[k + "mu_delta"]
        out = pb.solve_iebpx(d["b"], 1e-9, nc, mu, delta)
        assert out["iters"] == int(g[k + "iters"])
        assert abs(out["ext_relres"] - g[k + "norms"][0]) <= 1e-12 * g[k + "norms"][0]
        assert abs(out["relres"] - g[k + "norms"][1]) <= 1e-12 * g[k + "norms"][1]
        assert np.max(np.abs(out["x"] - g[k + "x"])) <= 1e-13 * np.max(np.abs(g[k + "x"]))
    assert out["relres"] < 1e-9                      # the extended-system iterate gives a solution of A x = f


def test_iebpx_matches_live_reference():
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.laplacian("7pt", 10)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    h.build_transfers(H.BPX, 0.7)
    pb = O.Problem(h, H.BPX, H.JACOBI, 0.7)
    lo, hi = pb.eigs_power(20)
    mu, delta = (hi + lo) / (hi - lo), 2.0 / (hi + lo)
    for nc in (1, 2, 3, 25, 400):
        want = O.RefSolver(h, H.IMPLICIT_EXTENDED_SYSTEM_BPX, H.JACOBI, b, 0.7, one_thread_per_level=True)
        r = want.solve_iebpx(nc, 1e-9, mu, delta)
        want.close()
        out = pb.solve_iebpx(b, 1e-9, nc, mu, delta)
        # thread 0 of the reference sums the threads' residual contributions without a barrier
        # (src/SMEM_ExtendedSystem.cpp:641-648): a stale contribution delays the tolerance stop by one iteration
        assert r["iters"] in (out["iters"], out["iters"] + 1)
        if r["iters"] == out["iters"]:
            assert np.max(np.abs(out["x"] - r["x"])) <= 1e-13 * np.max(np.abs(r["x"]))
            assert abs(out["ext_relres"] - r["ext_relres"]) <= 1e-13 + 1e-9 * r["ext_relres"]


# ---- explicit extended-system BPX (`-solver eebpx`): the same loop on the assembled extended matrix -----------------
def _explicit_oracle(h, AA, disp, bb, b, nc, mu, delta):
    h1 = H.Hierarchy([AA], [])
    h1.P, h1.R = [], []
    out = O.Problem(h1, H.BPX, H.JACOBI, 1.0).solve_iebpx(bb, 1e-9, nc, mu, delta)
    x = H.extended_solution(h, disp, out["x"])
    rel = O.norm2(O.spgemv(h.A[0], x, b, -1.0, 1.0)) / O.norm2(b)
    return dict(x=x, xx=out["x"], iters=out["iters"], ext_relres=out["relres"], relres=rel)


def test_extended_matrix_blocks_and_equivalence_with_the_implicit_form():
    """BuildExtendedMatrix restated on the host: block (k,l) = A_k P_k..P_{l-1}, symmetric, diag first; Jacobi-Chebyshev on it
    is the same iteration as the implicit form with weight 1"""
    A = H.laplacian("7pt", 9)
    h = H.amg_setup(A)
    h.build_transfers(H.BPX, 1.0)
    b = H.rand_rhs(A.nrows)
    AA, disp, bb = H.extended_system(h, b)
    assert np.array_equal(AA.indices[AA.indptr[:-1]], np.arange(AA.nrows))
    S = AA.to_scipy().copy()          # (scipy sorts a row's indices in place when slicing: keep AA's diag-first rows intact)
    As, Ps = [a.to_scipy().copy() for a in h.A], [p.to_scipy().copy() for p in h.P]
    for k in range(h.num_levels):
        M = As[k]
        for l in range(k, h.num_levels):
            if l > k:
                M = M @ Ps[l - 1]
            assert abs(S[disp[k]:disp[k + 1], disp[l]:disp[l + 1]] - M).max() <= 1e-13 * abs(S).max()
            assert abs(S[disp[l]:disp[l + 1], disp[k]:disp[k + 1]] - M.T).max() <= 1e-13 * abs(S).max()
    pb = O.Problem(h, H.BPX, H.JACOBI, 1.0)
    lo, hi = pb.eigs_power(20)
    mu, delta = (hi + lo) / (hi - lo), 2.0 / (hi + lo)
    imp = pb.solve_iebpx(b, 1e-9, 300, mu, delta)
    exp = _explicit_oracle(h, AA, disp, bb, b, 300, mu, delta)
    assert imp["iters"] == exp["iters"] and imp["relres"] < 1e-9
    assert np.max(np.abs(imp["x"] - exp["x"])) <= 1e-12 * np.max(np.abs(imp["x"]))


def test_extended_matrix_assembly_matches_live_reference():
    """BuildExtendedMatrix (src/SMEM_Setup.cpp:1426-1521) of the reference's object code (hypre_CSRMatrixMultiply / Transpose are
    single-rank stand-ins following hypre's published algorithms) against hierarchy.extended_system: the same operator to 1e-14,
    the same block offsets, row lengths and diagonal-first rows; inside a product block the entry order differs (hypre keeps
    first-touch order, the host product sorts), which only moves the summation order of the SpMV"""
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    for prob, n in (("7pt", 9), ("5pt", 20)):
        A = H.laplacian(prob, n)
        h = H.amg_setup(A)
        h.build_transfers(H.BPX, 1.0)
        AA, disp, bb = H.extended_system(h, H.rand_rhs(A.nrows))
        R, d2 = O.ref_build_extended_matrix(h)
        assert list(disp) == list(d2) and (AA.nrows, AA.nnz) == (R.nrows, R.nnz)
        assert np.array_equal(AA.indptr, R.indptr)
        assert np.array_equal(R.indices[R.indptr[:-1]], np.arange(R.nrows))                 # diagonal first
        assert abs(AA.to_scipy() - R.to_scipy()).max() <= 1e-14 * abs(R.to_scipy()).max()
        for r in range(0, AA.nrows, 37):                                                    # same column sets row by row
            assert set(AA.indices[AA.indptr[r]:AA.indptr[r + 1]]) == set(R.indices[R.indptr[r]:R.indptr[r + 1]])


@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
def test_eebpx_matches_reference_fixture(name):
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "iebpx.npz")))
    h, d = hierarchy_from_golden(name)
    h.build_transfers(H.BPX, 1.0)
    AA, disp, bb = H.extended_system(h, d["b"])
    for nc in (2, 7, 300):
        k = "%s_explicit_nc%d_" % (name, nc)
        mu, delta = g[k + "mu_delta"]
        out = _explicit_oracle(h, AA, disp, bb, d["b"], nc, mu, delta)
        assert out["iters"] == int(g[k + "iters"])
        # (norms of a converged iterate carry the cancellation of f - A x: absolute 1e-13 of the initial residual)
        assert abs(out["ext_relres"] - g[k + "norms"][0]) <= 1e-13 + 1e-9 * g[k + "norms"][0]
        assert abs(out["relres"] - g[k + "norms"][1]) <= 1e-13 + 1e-9 * g[k + "norms"][1]
        assert np.max(np.abs(out["x"] - g[k + "x"])) <= 1e-12 * np.max(np.abs(g[k + "x"]))


def test_eebpx_matches_live_reference():
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.laplacian("5pt", 20)
    h = H.amg_setup(A)
    h.build_transfers(H.BPX, 1.0)
    b = H.rand_rhs(A.nrows)
    AA, disp, bb = H.extended_system(h, b)
    lo, hi = O.Problem(h, H.BPX, H.JACOBI, 1.0).eigs_power(20)
    mu, delta = (hi + lo) / (hi - lo), 2.0 / (hi + lo)
    for nc, nt in ((1, 1), (2, 4), (9, 3), (300, 4)):
        r = O.ref_solve_eebpx(h, AA, disp, bb, nc, 1e-9, mu, delta, num_threads=nt)
        out = _explicit_oracle(h, AA, disp, bb, b, nc, mu, delta)
        assert r["iters"] in (out["iters"], out["iters"] + 1)       # (the same unsynchronised sum, :641-648)
        if r["iters"] == out["iters"]:
            assert np.max(np.abs(out["xx"] - r["xx"])) <= 1e-13 * np.max(np.abs(r["xx"]))
            assert np.max(np.abs(out["x"] - r["x"])) <= 1e-13 * np.max(np.abs(r["x"]))


# ---- nonsymmetric operators: `-problem difconv` (src/BuildHypreMatrix.cpp:104-245) ---------------------------------------
@pytest.mark.parametrize("a,atype", [((40.0, -20.0, 10.0), 3), ((10.0, 10.0, 10.0), 0), ((5.0, 5.0, 5.0), -1)])
def test_nonsymmetric_difconv_history_matches_live_reference(a, atype):
    """convection-diffusion stencils (upwind / forward / centred): the oracle's synchronous Multadd history against the
    reference's object code on a NONSYMMETRIC hierarchy (R-bar = P^T GT scales A's columns, src/SMEM_Setup.cpp:1209-1253)"""
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.difconv(12, a=a, atype=atype)
    S = A.to_scipy().copy()
    assert abs(S - S.T).max() > 1.0                                  # really nonsymmetric
    assert abs(S[12 * 12 * 5 + 12 * 5 + 5].sum()) < 1e-10 * abs(S).max()   # interior row sums vanish (pure derivatives)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    h.build_transfers(H.MULTADD, 0.9)
    _, hist, _ = O.Problem(h, H.MULTADD, H.JACOBI, 0.9).solve_sync(b, 1e-9, 100)
    rs = O.RefSolver(h, H.MULTADD, H.JACOBI, b, 0.9, one_thread_per_level=True)
    out = rs.solve_sync_det(100, 1e-9)
    rs.close()
    assert len(hist) == len(out["hist"]) and hist[-1] < 1e-9
    assert np.max(np.abs(hist - out["hist"])) <= HIST_TOL


# ---- hybrid Jacobi / Gauss-Seidel pinned by the reference's object code (block = a thread's row range) ---------------------
@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
@pytest.mark.parametrize("tag", ["nt0", "nt16"])
def test_hybrid_jgs_history_matches_reference_fixture(name, tag):
    """SMEM_Sync_HybridJacobiGaussSeidel inside Multadd (-num_post_smooth_sweeps 0) with one and with several threads per
    level: the oracle replays the reference's blocks (the nnz-balanced thread ranges, src/SMEM_Setup.cpp:945-979)"""
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "hybrid_jgs.npz")))
    h, d = hierarchy_from_golden(name)
    h.build_transfers(H.MULTADD, 0.9, num_pre=1, num_post=0)
    tpl = g["%s_%s_threads_per_level" % (name, tag)]
    blocks = [H.nnz_balanced_bounds(h.A[l].indptr, int(tpl[l])) for l in range(h.num_levels)]
    _, hist, _ = O.Problem(h, H.MULTADD, H.HYBRID_JACOBI_GAUSS_SEIDEL, 0.9, num_pre=1, num_post=0,
                           jgs_blocks=blocks).solve_sync(d["b"], 1e-9, 80)
    _close_hist(hist, g["%s_%s_hist" % (name, tag)])
    assert hist[-1] < 1e-9


@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
def test_chebyshev_accelerated_bpx_matches_reference_fixture(name):
    """the -cheby branch of SMEM_Solve (src/SMEM_Solve.cpp:169-188) around the BPX cycle, from the reference's object code"""
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "cheby_bpx.npz")))
    h, d = hierarchy_from_golden(name)
    h.build_transfers(H.BPX, 0.8)
    mu, delta = g[name + "_mu_delta"]
    _, hist, _ = O.Problem(h, H.BPX, H.JACOBI, 0.8).solve_sync(d["b"], 1e-9, 200, cheby=(mu, delta))
    _close_hist(hist, g[name + "_hist"])
    assert hist[-1] < 1e-9


def test_transpose_matvec_matches_live_reference():
    """-no_construct_R: SMEM_Sync_Parfor_MatVecT of the reference's object code = the oracle's scatter form
    (bit-identical with one thread; per-thread partial sums otherwise) = the explicit R = P^T product the default path uses"""
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.laplacian("7pt", 11)
    h = H.amg_setup(A)
    h.build_transfers(H.AFACX, 0.9)                      # plain P, R = P^T
    rng = np.random.default_rng(11)
    for l in range(h.num_levels - 1):
        x = rng.uniform(-1, 1, h.n[l])
        seq = O.matvecT(h.P[l], x)
        assert np.array_equal(O.ref_parfor_matvec_t(h.P[l], x, 1), seq)
        mag = np.maximum(O.matvecT(H.CSR(h.P[l].nrows, h.P[l].ncols, h.P[l].indptr, h.P[l].indices, np.abs(h.P[l].data)), np.abs(x)), 1e-300)
        for nt in (3, 8):
            assert np.max(np.abs(O.ref_parfor_matvec_t(h.P[l], x, nt) - seq) / mag) <= 4e-16
        assert np.max(np.abs(O.spgemv(h.R[l], x, None, 1.0, 0.0) - seq) / mag) <= 4e-16


@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
def test_par_bpx_matches_reference_fixture(name):
    """`-solver par_bpx` from the reference's object code (tests/golden/par_bpx.npz): BPX with the Jacobi weight applied twice
    (xx = w rr / (a_ii / w), src/SMEM_Sync_AMG.cpp:213-218 with src/SMEM_Setup.cpp:451-460), and with the step w / l1 for L1"""
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "par_bpx.npz")))
    h, d = hierarchy_from_golden(name)
    h.build_transfers(H.BPX, 0.6)
    sv, sm, w2 = H.par_bpx_equivalent(H.JACOBI, 0.6)
    assert (sv, sm) == (H.BPX, H.JACOBI) and abs(w2 - 0.36) < 1e-15
    u, hist, _ = O.Problem(h, sv, sm, w2).solve_sync(d["b"], 1e-30, 12)
    want = g[name + "_j_hist"]
    assert len(hist) == len(want) and np.max(np.abs(hist - want) / want) <= 1e-10
    assert np.max(np.abs(u - g[name + "_j_u"])) <= 1e-12 * np.max(np.abs(u))
    u, hist, _ = O.Problem(h, H.BPX, H.L1_JACOBI, 0.6, l1_scale=1.0 / 0.6).solve_sync(d["b"], 1e-30, 12)
    want = g[name + "_l1_hist"]
    assert len(hist) == len(want) and np.max(np.abs(hist - want) / want) <= 1e-10
    with pytest.raises(ValueError):
        H.par_bpx_equivalent(H.L1_JACOBI, 0.6)


def test_parfor_afacx_cycle_matches_live_reference():
    """SMEM_Sync_Parfor_AFACx_Vcycle (src/SMEM_Sync_AMG.cpp:296-406): the omp-for form of AFACx, which the
    reference's driver never selects (AFACX always gets the ALL_LEVELS partition, src/SMEM_Main.cpp:641-649) -- reached here by
    forcing ONE_LEVEL in the object-code driver; 4 threads, deterministic; equals the oracle's AFACx history"""
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.laplacian("7pt", 10)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    h.build_transfers(H.AFACX, 0.6)
    O.ref_lib().ref_force_one_level(1)
    try:
        for sweeps in (1, 2):
            rs = O.RefSolver(h, H.AFACX, H.JACOBI, b, 0.6, num_threads=4, fine_sweeps=sweeps, coarse_sweeps=sweeps)
            out = rs.solve(40, 1e-9, async_type=0)
            rs.close()
            _, hist, _ = O.Problem(h, H.AFACX, H.JACOBI, 0.6, fine_sweeps=sweeps, coarse_sweeps=sweeps).solve_sync(b, 1e-9, 40)
            _close_hist(hist, out["hist"])
    finally:
        O.ref_lib().ref_force_one_level(0)


def test_async_gauss_seidel_smoothers_match_live_reference():
    """SMEM_Async_GaussSeidel / SMEM_SemiAsync_GaussSeidel (src/SMEM_Smooth.cpp:445-502) inside Multadd, the
    reference's object code with one thread per level: the chaotic sweep over a single row range IS Gauss-Seidel = the oracle's
    hybrid smoother with one block"""
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.laplacian("7pt", 10)
    h = H.amg_setup(A)
    if h.num_levels > O.ref_lib().ref_max_threads():
        pytest.skip("more levels than cores: the reference's spin barriers would oversubscribe")
    b = H.rand_rhs(A.nrows)
    h.build_transfers(H.MULTADD, 0.9, num_pre=1, num_post=0)
    blocks = [np.asarray([0, a.nrows], dtype=np.int32) for a in h.A]
    for sm in (H.ASYNC_GAUSS_SEIDEL, H.SEMI_ASYNC_GAUSS_SEIDEL):
        for sweeps in (1, 2):
            _, want, _ = O.Problem(h, H.MULTADD, H.HYBRID_JACOBI_GAUSS_SEIDEL, 0.9, num_pre=1, num_post=0, jgs_blocks=blocks,
                                   fine_sweeps=sweeps).solve_sync(b, 1e-9, 60)
            rs = O.RefSolver(h, H.MULTADD, sm, b, 0.9, num_pre=1, num_post=0, one_thread_per_level=True, fine_sweeps=sweeps)
            out = rs.solve_sync_det(60, 1e-9)
            rs.close()
            _close_hist(want, out["hist"])
            assert want[-1] < 1e-9


# ---- SMEM_Async_Add_AMG (src/SMEM_Async_AMG.cpp), SURVEY.md row a15 -----------------------------------------------------------
_ASYNC_CASES = (("multadd", H.ASYNC_MULTADD, H.MULTADD, 0.9, 1), ("afacx", H.ASYNC_AFACX, H.AFACX, 0.6, 1), ("afacx2", H.ASYNC_AFACX, H.AFACX, 0.6, 2))


@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
def test_async_single_group_matches_reference_fixture(name):
    """the asynchronous solver of the reference's object code on a two-level hierarchy (one working group: deterministic)
    against the oracle's sequential model, Multadd and AFACx chains (tests/golden/async_two_level.npz)"""
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "async_two_level.npz")))
    hf, d = hierarchy_from_golden(name)
    h = H.Hierarchy(hf.A[:2], hf.P_plain[:1])
    for tag, solver, base, w, sweeps in _ASYNC_CASES:
        h.build_transfers(base, w)
        pb = O.Problem(h, base, H.JACOBI, w, fine_sweeps=sweeps, coarse_sweeps=sweeps)
        for K in (1, 7, 30):
            u, counts, rel = pb.solve_async_sequential(d["b"], K)
            want = g["%s_%s_k%d_u" % (name, tag, K)]
            assert list(counts) == [K, K]
            assert np.max(np.abs(u - want)) <= 1e-14 * np.max(np.abs(want)), (tag, K)
            assert abs(rel - float(g["%s_%s_k%d_relres" % (name, tag, K)])) <= 1e-13


def test_async_single_group_matches_live_reference():
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    if O.ref_lib().ref_max_threads() < 2:
        pytest.skip("needs two cores (one spinning thread per level)")
    A = H.laplacian("27pt", 8)
    h = H.amg_setup(A, max_levels=2)
    b = H.rand_rhs(A.nrows)
    for tag, solver, base, w, sweeps in _ASYNC_CASES:
        h.build_transfers(base, w)
        for K in (2, 11):
            rs = O.RefSolver(h, solver, H.JACOBI, b, w, one_thread_per_level=True, fine_sweeps=sweeps, coarse_sweeps=sweeps)
            out = rs.solve(K, 1e-9, async_type=0)
            rs.close()
            u, counts, rel = O.Problem(h, base, H.JACOBI, w, fine_sweeps=sweeps, coarse_sweeps=sweeps).solve_async_sequential(b, K)
            assert list(out["corrections"]) == list(counts) == [K, K]
            assert np.max(np.abs(u - out["u"])) <= 1e-14 * np.max(np.abs(u)), (tag, K)
            assert abs(rel - out["relres"]) <= 1e-13


def test_async_read_res_mode_matches_live_reference():
    """`-read_type res` of SMEM_Async_Add_AMG (src/SMEM_Async_AMG.cpp:227-236,285-296,416-426): the shared residual is updated
    incrementally (r -= A_0 e) and u is assembled from the per-level accumulators at the end -- bit for bit on two levels"""
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    if O.ref_lib().ref_max_threads() < 2:
        pytest.skip("needs two cores (one spinning thread per level)")
    A = H.laplacian("7pt", 10)
    h = H.amg_setup(A, max_levels=2)
    b = H.rand_rhs(A.nrows)
    O.ref_lib().ref_set_read_type(1)
    try:
        for tag, solver, base, w, sweeps in _ASYNC_CASES:
            h.build_transfers(base, w)
            for K in (1, 3, 20):
                rs = O.RefSolver(h, solver, H.JACOBI, b, w, one_thread_per_level=True, fine_sweeps=sweeps, coarse_sweeps=sweeps)
                out = rs.solve(K, 1e-9, async_type=0)
                rs.close()
                pb = O.Problem(h, base, H.JACOBI, w, fine_sweeps=sweeps, coarse_sweeps=sweeps)
                u, counts, rel = pb.solve_async_sequential(b, K, read_res=True)
                assert list(out["corrections"]) == list(counts) == [K, K]
                # The reference assembles u after the loop with `omp for` over the levels' accumulators WITHOUT a barrier in
                # front (:416-426): the idle coarsest-level thread gets there first and adds level 0's accumulator for ITS
                # half of the rows while level 0 is still correcting -- those rows can miss corrections (a race of the
                # reference; the oracle follows the race-free meaning).  Thread 0's rows (static schedule: the first half) are
                # added by the working thread itself after its last correction and are always complete.
                half = h.n[0] // 2
                assert np.max(np.abs(u[:half] - out["u"][:half])) <= 1e-14 * np.max(np.abs(u)), (tag, K)
                if np.max(np.abs(u - out["u"])) <= 1e-14 * np.max(np.abs(u)):
                    assert abs(rel - out["relres"]) <= 1e-13
                u2, _, _ = pb.solve_async_sequential(b, K)          # READ_SOL: the same iterate up to rounding
                assert np.max(np.abs(u - u2)) <= 1e-12 * np.max(np.abs(u))
    finally:
        O.ref_lib().ref_set_read_type(0)


# ---- ChebySetup / EigsPower / BPXCycle (src/SMEM_Cheby.cpp), SURVEY.md row a17 ------------------------------------------------
_CHEBY_SETUP_CASES = (("j", H.JACOBI, 0.8), ("l1", H.L1_JACOBI, 0.8), ("hjgs", H.HYBRID_JACOBI_GAUSS_SEIDEL, 1.0))


def _oracle_cheby_setup(h, sm, w, iters, nthreads):
    blocks = [H.nnz_balanced_bounds(a.indptr, nthreads) for a in h.A]       # BPXCycle's block = hypre's load-balanced range
    lo, hi = O.Problem(h, H.BPX, sm, w, jgs_blocks=blocks).eigs_power(iters)
    return np.asarray([lo, hi, (hi + lo) / (hi - lo), 2.0 / (hi + lo)])


@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
def test_cheby_setup_matches_reference_fixture(name):
    """orc_eigs_power + the mu / delta formulas against ChebySetup -> EigsPower -> BPXCycle of the reference's object code
    (tests/golden/cheby_setup.npz): eigenvalue bounds and Chebyshev scalars to 1e-12 relative"""
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "cheby_setup.npz")))
    h, d = hierarchy_from_golden(name)
    h.build_transfers(H.BPX, 0.8)
    for tag, sm, w in _CHEBY_SETUP_CASES:
        for iters in (3, 20):
            want = g["%s_%s_it%d" % (name, tag, iters)]
            got = _oracle_cheby_setup(h, sm, w, iters, 4)
            assert np.max(np.abs(got - want) / np.abs(want)) <= 1e-12, (tag, iters, got, want)


def test_cheby_setup_matches_live_reference():
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    for prob, n in (("7pt", 9), ("27pt", 7)):
        A = H.laplacian(prob, n)
        h = H.amg_setup(A)
        b = H.rand_rhs(A.nrows)
        h.build_transfers(H.BPX, 0.7)
        for tag, sm, w in _CHEBY_SETUP_CASES:
            for iters, nt in ((2, 1), (7, 3), (20, 4)):
                r = O.ref_cheby_setup(h, b, sm, w, iters, 1, nt)
                assert np.array_equal(r["f_after"], b)                       # EigsPower restores the right-hand side
                want = np.asarray([r["alpha"], r["beta"], r["mu"], r["delta"]])
                got = _oracle_cheby_setup(h, sm, w, iters, nt)
                assert np.max(np.abs(got - want) / np.abs(want)) <= 1e-12, (prob, tag, iters, nt)


def test_chebyshev_accelerated_vcycle_matches_live_reference():
    """-cheby around the multiplicative V-cycle (precond_flag = 1 form of SMEM_Sync_Parfor_Vcycle, src/SMEM_Sync_AMG.cpp:29-36:
    the cycle on the residual from a zero guess), the reference's object code with 4 threads (omp-for: deterministic).  The device
    rejects this combination (Chebyshev acceleration is wired for the additive cycles); the oracle restates it."""
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.laplacian("7pt", 10)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    h.build_transfers(H.MULT, 0.8)
    for lo, hi in ((0.5, 1.0), (0.3, 1.2)):
        mu, delta = (hi + lo) / (hi - lo), 2.0 / (hi + lo)
        rs = O.RefSolver(h, H.MULT, H.JACOBI, b, 0.8, num_threads=4)
        out = rs.solve(100, 1e-9, async_type=0, cheby=(mu, delta), precond=1)
        rs.close()
        _, hist, _ = O.Problem(h, H.MULT, H.JACOBI, 0.8).solve_sync(b, 1e-9, 100, cheby=(mu, delta))
        _close_hist(hist, out["hist"])
        assert hist[-1] < 1e-9


# ---- DMEM: synchronous Multadd on all ranks and the acceleration of the accumulated correction -----------------------------
@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
def test_dmem_sync_add_matches_reference_fixture(name):
    """DMEM_SyncAdd / DMEM_SyncAddCycle (src/DMEM_Mult.cpp:263-450; direct solve on the coarsest level, symmetrised and plain
    smoother) and DMEM_ChebyUpdate (src/DMEM_Misc.cpp:612-666; Richardson and Chebyshev branches) from the reference's object
    code compiled for one rank (tests/golden/dmem.npz) against the oracle's restatement"""
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "dmem.npz")))
    h, d = hierarchy_from_golden(name)
    b, w = d["b"], 0.9
    for tag, post in (("sym", 1), ("plain", 0)):
        h.build_transfers(H.MULTADD, w, num_pre=1, num_post=post)
        u, hist = O.Problem(h, H.MULTADD, H.JACOBI, w, num_pre=1, num_post=post, coarse_solve=1).solve_sync_dmem(b, 1e-9, 100)
        _close_hist(hist, g["%s_%s_hist" % (name, tag)])
        assert np.max(np.abs(u - g["%s_%s_x" % (name, tag)])) <= 1e-12 * np.max(np.abs(u))
    h.build_transfers(H.MULTADD, w)
    pb = O.Problem(h, H.MULTADD, H.JACOBI, w, coarse_solve=1)
    mu, delta = g[name + "_mu_delta"]
    _, hist = pb.solve_sync_dmem(b, 1e-9, 100, 2, mu, delta)             # ours: 2 = second-order Richardson
    _close_hist(hist, g[name + "_richardson_hist"])
    _, hist = pb.solve_sync_dmem(b, 1e-9, 100, 1, mu, delta)             # ours: 1 = Chebyshev recurrence
    _close_hist(hist, g[name + "_chebyshev_hist"])


@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
def test_dmem_addcycle_matches_reference_fixture(name):
    """AddCycle (src/DMEM_Add.cpp:180-329) + DMEM_AddSmooth (src/DMEM_Smooth.cpp:574-638) from the reference's object code, every
    grid in turn on one rank = the oracle's sequential model of the asynchronous solve with the DMEM coarse solve"""
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "dmem.npz")))
    h, d = hierarchy_from_golden(name)
    h.build_transfers(H.MULTADD, 0.9)
    u, counts, rel = O.Problem(h, H.MULTADD, H.JACOBI, 0.9, coarse_solve=1).solve_async_sequential(d["b"], 12)
    assert abs(rel - g[name + "_addcycle_hist"][-1]) <= HIST_TOL and list(counts) == [12] * h.num_levels
    assert np.max(np.abs(u - g[name + "_addcycle_x"])) <= 1e-12 * np.max(np.abs(u))


@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
def test_dmem_async_smooth_matches_reference_fixture(name):
    """DMEM_AsyncSmooth (src/DMEM_Smooth.cpp:16-313) from the reference's object code on one rank: u = r ./ s, x += u,
    r -= A_diag u, `num_cycles` relaxations -- exactly that many (L1-)Jacobi sweeps of the oracle"""
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "dmem.npz")))
    h, d = hierarchy_from_golden(name)
    want = O.smooth("jacobi", h.A[0], d["b"], 0.9, sweeps=17, zero_flag=1)
    assert np.max(np.abs(want - g[name + "_asyncsmooth_j_x"])) <= 1e-13 * np.max(np.abs(want))
    want = O.smooth("l1_jacobi", h.A[0], d["b"], 0.9, sweeps=17, zero_flag=1, l1=h.l1_norms()[0])
    assert np.max(np.abs(want - g[name + "_asyncsmooth_l1_x"])) <= 1e-13 * np.max(np.abs(want))


@pytest.mark.parametrize("name", ["lap5pt_n32", "lap7pt_n12"])
def test_dmem_mult_matches_reference_fixture(name):
    """DMEM_Mult / DMEM_MultCycle (src/DMEM_Mult.cpp:13-261), the DMEM driver's multiplicative comparator, from the reference's
    object code on one rank (tests/golden/dmem_mult.npz) against the oracle's V-cycle with the direct coarse solve"""
    import os
    from conftest import GOLDEN, hierarchy_from_golden
    g = dict(np.load(os.path.join(GOLDEN, "dmem_mult.npz")))
    h, d = hierarchy_from_golden(name)
    h.build_transfers(H.MULT, 0.8)
    u, hist, _ = O.Problem(h, H.MULT, H.JACOBI, 0.8, coarse_solve=1).solve_sync(d["b"], 1e-9, 100)
    _close_hist(hist, g[name + "_hist"])
    assert hist[-1] < 1e-9
    assert np.max(np.abs(u - g[name + "_x"])) <= 1e-12 * np.max(np.abs(u))


def test_dmem_mult_matches_live_reference():
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.laplacian("27pt", 8)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    for w in (0.6, 0.9):
        h.build_transfers(H.MULT, w)
        x, want = O.ref_dmem_mult(h, b, w, 80, 1e-9)
        u, hist, _ = O.Problem(h, H.MULT, H.JACOBI, w, coarse_solve=1).solve_sync(b, 1e-9, 80)
        _close_hist(hist, want)
        assert np.max(np.abs(u - x)) <= 1e-12 * np.max(np.abs(u))


def test_dmem_sync_bpx_mapping_matches_live_reference():
    """DMEM's SYNC_BPX (src/DMEM_Mult.cpp:346-349) is DMEM_SyncAddCycle on the PLAIN interpolants with the direct coarse solve: the
    reference's object code fed plain P equals the oracle's Multadd with plain transfers, plain Jacobi and coarse_solve = 1"""
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.laplacian("7pt", 10)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    h.build_transfers(H.BPX, 0.5)
    x, want = O.ref_dmem_sync_add(h, b, 0.5, symmetrised=False, num_cycles=30, tol=1e-9)
    u, hist = O.Problem(h, H.MULTADD, H.JACOBI, 0.5, num_pre=1, num_post=0, coarse_solve=1).solve_sync_dmem(b, 1e-9, 30)
    assert len(hist) == len(want) and np.max(np.abs(hist - want) / want) <= 1e-10      # (not a convergent iteration without acceleration)
    assert np.max(np.abs(u - x)) <= 1e-10 * np.max(np.abs(x))


def test_dmem_sync_add_matches_live_reference():
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.laplacian("7pt", 11)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    for sym, post in ((True, 1), (False, 0)):
        h.build_transfers(H.MULTADD, 0.8, num_pre=1, num_post=post)
        u, hist = O.Problem(h, H.MULTADD, H.JACOBI, 0.8, num_pre=1, num_post=post, coarse_solve=1).solve_sync_dmem(b, 1e-9, 100)
        x, rh = O.ref_dmem_sync_add(h, b, 0.8, symmetrised=sym, num_cycles=100, tol=1e-9)
        _close_hist(hist, rh)
        assert np.max(np.abs(u - x)) <= 1e-12 * np.max(np.abs(x))


def test_l1_hybrid_jgs_parfor_smoother_matches_live_reference():
    """L1_HYBRID_JACOBI_GAUSS_SEIDEL (smoother 12): reachable through the Parfor smoother only (BPX / ONE_LEVEL partition,
    src/SMEM_Solve.cpp:324-334), where SMEM_Sync_Parfor_HybridJacobiGaussSeidel divides by hypre's l1 norms instead of a_ii / w
    (src/SMEM_Smooth.cpp:253-263).  One thread = one Gauss-Seidel block per level; BPX is run unaccelerated for a few cycles
    (it need not converge -- the histories must agree)."""
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    A = H.laplacian("7pt", 9)
    h = H.amg_setup(A)
    h.build_transfers(H.BPX, 1.0)
    b = H.rand_rhs(A.nrows)
    for sm in (H.L1_HYBRID_JACOBI_GAUSS_SEIDEL, H.HYBRID_JACOBI_GAUSS_SEIDEL):
        rs = O.RefSolver(h, H.BPX, sm, b, 0.8, num_threads=1)
        out = rs.solve(6, 1e-300, async_type=0)
        rs.close()
        _, hist, _ = O.Problem(h, H.BPX, sm, 0.8).solve_sync(b, 1e-300, 6)
        assert len(hist) == 7 and np.all(np.isfinite(hist))
        assert np.max(np.abs(hist - out["hist"]) / np.maximum(np.abs(out["hist"]), 1e-300)) <= 1e-12
