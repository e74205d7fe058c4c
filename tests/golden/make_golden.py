"""Generates tests/golden/*.npz: small hierarchies together with outputs of the REFERENCE's own
object code (oracle/_ref/libref_smem.so, built from /root/reference/src by oracle/build_ref.sh).
Run in the build container (needs /root/reference); the fixtures are committed so that the GPU
box -- which has no /root/reference -- can check the oracle and the CUDA path against them.

    python tests/golden/make_golden.py            (everything; `--<name>-only` regenerates one file)

A regeneration moves the histories in the last bits (1e-16: the reference's threaded norm reductions are not run-to-run
deterministic); the single-thread / one-thread-per-level outputs (async_two_level, cheby_setup, smooth_transfer, dmem_mult) come
back byte-identical.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import async_multigrid_b200 as amg  # noqa: E402
from async_multigrid_b200 import hierarchy as H  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def pack(h):
    d = {"num_levels": np.asarray(h.num_levels)}
    for l in range(h.num_levels):
        for nm, m in (("A", h.A[l]),) + ((("Pp", h.P_plain[l]),) if l < h.num_levels - 1 else ()):
            d["%s%d_shape" % (nm, l)] = np.asarray(m.shape)
            d["%s%d_indptr" % (nm, l)] = m.indptr
            d["%s%d_indices" % (nm, l)] = m.indices
            d["%s%d_data" % (nm, l)] = m.data
    return d


def main():
    from oracle import build as obuild
    amg.build.build_host(); obuild.build_oracle(); obuild.build_ref()
    assert O.ref_lib() is not None, "oracle/_ref not available"
    for name, prob, n, w in (("lap5pt_n32", "5pt", 32, 0.9), ("lap7pt_n12", "7pt", 12, 0.9)):
        A = H.laplacian(prob, n)
        h = H.amg_setup(A)
        b = H.rand_rhs(A.nrows)
        d = pack(h)
        d["b"] = b
        d["smooth_weight"] = np.asarray(w)
        # reference: synchronous Multadd, symmetrised Jacobi (defaults pre=post=1), race-free run
        h.build_transfers(H.MULTADD, w)
        rs = O.RefSolver(h, H.MULTADD, H.JACOBI, b, w, one_thread_per_level=True)
        out = rs.solve_sync_det(100, 1e-9)
        d["multadd_symj_hist"] = out["hist"]
        d["multadd_symj_u"] = out["u"]
        rs.close()
        # reference: Multadd with plain Jacobi (post sweeps = 0: only R is smoothed)
        h.build_transfers(H.MULTADD, w, num_pre=1, num_post=0)
        rs = O.RefSolver(h, H.MULTADD, H.JACOBI, b, w, num_pre=1, num_post=0, one_thread_per_level=True)
        out = rs.solve_sync_det(60, 1e-9)
        d["multadd_j_hist"] = out["hist"]
        rs.close()
        # reference: L1-Jacobi symmetrised
        h.build_transfers(H.MULTADD, w, smooth_interp_type=H.L1_JACOBI)
        rs = O.RefSolver(h, H.MULTADD, H.L1_JACOBI, b, w, one_thread_per_level=True)
        out = rs.solve_sync_det(100, 1e-9)
        d["multadd_syml1_hist"] = out["hist"]
        rs.close()
        # reference: AFACx (grouped cycle), weight 0.6
        h.build_transfers(H.AFACX, 0.6)
        rs = O.RefSolver(h, H.AFACX, H.JACOBI, b, 0.6, one_thread_per_level=True)
        out = rs.solve_sync_det(40, 1e-9)
        d["afacx_j_hist"] = out["hist"]
        rs.close()
        # reference: BPX through SMEM_Solve itself (omp-for cycle, deterministic), 4 threads, 10 cycles
        rs = O.RefSolver(h, H.BPX, H.JACOBI, b, 0.6, num_threads=4)
        out = rs.solve(10, 1e-30, async_type=0)
        d["bpx_j_hist"] = out["hist"]
        rs.close()
        # reference: multiplicative V(1,1) cycle through SMEM_Solve itself (omp-for cycle, deterministic), weight 0.8
        h.build_transfers(H.MULT, 0.8)
        rs = O.RefSolver(h, H.MULT, H.JACOBI, b, 0.8, num_threads=4)
        out = rs.solve(100, 1e-9, async_type=0)
        d["mult_j_hist"] = out["hist"]
        rs.close()
        # reference kernels: SMEM_MatVec and the sequential smoothers on level 0
        x = np.cos(np.arange(A.nrows) * 0.37)
        y = np.zeros(A.nrows)
        import ctypes as C
        s = O.c_csr(h.A[0])
        O.ref_lib().ref_matvec(C.byref(s), O.dptr(x), O.dptr(y))
        d["x"] = x
        d["matvec_A0_x"] = y
        u = np.zeros(A.nrows)
        O.ref_lib().ref_seq_symmetric_jacobi(C.byref(s), O.dptr(b.copy()), O.dptr(u), w, 1)
        d["seq_symj_b"] = u
        u = np.zeros(A.nrows)
        O.ref_lib().ref_seq_jacobi(C.byref(s), O.dptr(b.copy()), O.dptr(u), w, 3, 1)
        d["seq_j3_b"] = u
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if "hist" in k})


def matrix_file_golden():
    """A small file in the reference's binary triplet format and what the reference's own reader
    (ReadBinary_fread_HypreParCSR, src/Misc.cpp:800-915, compiled in oracle/_ref) makes of it, for both settings of
    symm_flag.  The records are shuffled so that the in-row order (file order, mirrored entries appended) is pinned."""
    import tempfile
    rng = np.random.default_rng(7)
    A = H.laplacian("7pt", 5, 4, 3).to_scipy().tocoo()
    vals = A.data * (1.0 + 0.01 * np.minimum(A.row, A.col)) + 0.001 * np.maximum(A.row, A.col)   # symmetric, distinct values
    keep = A.col <= A.row
    rec = np.zeros(1 + int(keep.sum()), dtype=[("i", "<i4"), ("j", "<i4"), ("v", "<f8")])
    rec[0] = (A.shape[0], A.shape[1], 0.0)
    order = rng.permutation(int(keep.sum()))
    rec["i"][1:] = A.row[keep][order] + 1
    rec["j"][1:] = A.col[keep][order] + 1
    rec["v"][1:] = vals[keep][order]
    d = {"file_bytes": np.frombuffer(rec.tobytes(), dtype=np.uint8)}
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as fp:
        fp.write(rec.tobytes())
        path = fp.name
    for flag in (0, 1):
        ip, ix, dv = O.ref_read_matrix(path, flag)
        d["symm%d_indptr" % flag], d["symm%d_indices" % flag], d["symm%d_data" % flag] = ip, ix, dv
    os.unlink(path)
    np.savez_compressed(os.path.join(OUT, "matrix_file.npz"), **d)
    print("matrix_file", rec.shape, d["symm1_indptr"][-1])


def iebpx_golden():
    """The reference's implicit extended-system BPX solver (SMEM_ExtendedSystemSolve, src/SMEM_ExtendedSystem.cpp,
    compiled in oracle/_ref) on the hierarchies of the committed fixtures, one thread per level (deterministic)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import hierarchy_from_golden
    d = {}
    for name in ("lap5pt_n32", "lap7pt_n12"):
        h, g = hierarchy_from_golden(name)
        b, w = g["b"], 0.8
        h.build_transfers(H.BPX, w)
        for sm, tag in ((H.JACOBI, "j"), (H.L1_JACOBI, "l1")):
            lo, hi = O.Problem(h, H.BPX, sm, w).eigs_power(20)
            mu, delta = (hi + lo) / (hi - lo), 2.0 / (hi + lo)
            for nc in (2, 7, 300):
                # thread 0 of the reference sums the threads' residual contributions without a barrier
                # (src/SMEM_ExtendedSystem.cpp:641-648); a stale one delays the tolerance stop by an iteration: keep the
                # race-free outcome = the smallest count of a few runs
                r = None
                for _ in range(5):
                    rs = O.RefSolver(h, H.IMPLICIT_EXTENDED_SYSTEM_BPX, sm, b, w, one_thread_per_level=True)
                    t = rs.solve_iebpx(nc, 1e-9, mu, delta)
                    rs.close()
                    if r is None or t["iters"] < r["iters"]:
                        r = t
                k = "%s_%s_nc%d_" % (name, tag, nc)
                d[k + "mu_delta"] = np.asarray([mu, delta])
                d[k + "iters"] = np.asarray(r["iters"])
                d[k + "norms"] = np.asarray([r["ext_relres"], r["relres"]])
                d[k + "x"] = r["x"]
                print(k, r["iters"], r["ext_relres"], r["relres"])
    # explicit form (`-solver eebpx`): the reference's EXPLICIT branch on the extended matrix assembled by the host side
    for name in ("lap5pt_n32", "lap7pt_n12"):
        h, g = hierarchy_from_golden(name)
        b = g["b"]
        h.build_transfers(H.BPX, 1.0)
        AA, disp, bb = H.extended_system(h, b)
        lo, hi = O.Problem(h, H.BPX, H.JACOBI, 1.0).eigs_power(20)
        mu, delta = (hi + lo) / (hi - lo), 2.0 / (hi + lo)
        for nc in (2, 7, 300):
            r = None
            for _ in range(5):
                t = O.ref_solve_eebpx(h, AA, disp, bb, nc, 1e-9, mu, delta, num_threads=4)
                if r is None or t["iters"] < r["iters"]:
                    r = t
            k = "%s_explicit_nc%d_" % (name, nc)
            d[k + "mu_delta"] = np.asarray([mu, delta])
            d[k + "iters"] = np.asarray(r["iters"])
            d[k + "norms"] = np.asarray([r["ext_relres"], r["relres"]])
            d[k + "x"] = r["x"]
            print(k, r["iters"], r["ext_relres"], r["relres"])
    np.savez_compressed(os.path.join(OUT, "iebpx.npz"), **d)


def hybrid_jgs_golden():
    """Multadd with the hybrid Jacobi / Gauss-Seidel smoother (SMEM_Sync_HybridJacobiGaussSeidel, src/SMEM_Smooth.cpp:533-586;
    -num_post_smooth_sweeps 0) through the reference's object code with SEVERAL threads per level: a thread's row range is
    its Gauss-Seidel block, so the block lists are stored with the history."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import hierarchy_from_golden
    d = {}
    for name in ("lap5pt_n32", "lap7pt_n12"):
        h, g = hierarchy_from_golden(name)
        b, w = g["b"], 0.9
        h.build_transfers(H.MULTADD, w, num_pre=1, num_post=0)
        for nt in (h.num_levels, 16):
            rs = O.RefSolver(h, H.MULTADD, H.HYBRID_JACOBI_GAUSS_SEIDEL, b, w, num_pre=1, num_post=0, num_threads=nt,
                             one_thread_per_level=(nt == h.num_levels))
            tpl = np.asarray(rs.threads_per_level, dtype=np.int32)
            out = rs.solve_sync_det(80, 1e-9)
            rs.close()
            k = "%s_nt%d_" % (name, 16 if nt == 16 else 0)
            d[k + "threads_per_level"] = tpl
            d[k + "hist"] = out["hist"]
            print(k, tpl, len(out["hist"]) - 1, out["hist"][-1])
    np.savez_compressed(os.path.join(OUT, "hybrid_jgs.npz"), **d)


def cheby_golden():
    """SMEM_Solve with -cheby (Chebyshev acceleration of the BPX cycle, src/SMEM_Solve.cpp:169-188, precond_flag = 1) through
    the reference's object code, 4 threads (omp-for cycle: deterministic)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import hierarchy_from_golden
    d = {}
    for name in ("lap5pt_n32", "lap7pt_n12"):
        h, g = hierarchy_from_golden(name)
        b, w = g["b"], 0.8
        h.build_transfers(H.BPX, w)
        lo, hi = O.Problem(h, H.BPX, H.JACOBI, w).eigs_power(20)
        mu, delta = (hi + lo) / (hi - lo), 2.0 / (hi + lo)
        rs = O.RefSolver(h, H.BPX, H.JACOBI, b, w, num_threads=4)
        out = rs.solve(200, 1e-9, async_type=0, cheby=(mu, delta), precond=1)
        rs.close()
        d[name + "_mu_delta"] = np.asarray([mu, delta])
        d[name + "_hist"] = out["hist"]
        print(name, len(out["hist"]) - 1, out["hist"][-1])
    np.savez_compressed(os.path.join(OUT, "cheby_bpx.npz"), **d)


def _dmem_accel_with_reference_update(h, pb, b, mu, delta, ref_accel_code, num_cycles=100, tol=1e-9):
    """DMEM_SyncAddCorrect's loop (src/DMEM_Add.cpp:706-711) with the reference's DMEM_ChebyUpdate object code applied to the
    cycle output"""
    S = h.A[0].to_scipy().copy()
    x, d, r = np.zeros_like(b), np.zeros_like(b), b.copy()
    c, cp, hist = mu, 1.0, [1.0]
    for k in range(num_cycles):
        e = pb.cycle(r)
        d, _, c, cp = O.ref_dmem_cheby_update(d, e, k, mu, delta, c, cp, accel_type=ref_accel_code)
        x = x + d
        r = b - S @ x
        hist.append(np.linalg.norm(r) / np.linalg.norm(b))
        if hist[-1] < tol:
            break
    return np.asarray(hist)


def dmem_golden():
    """DMEM_SyncAdd / DMEM_SyncAddCycle (src/DMEM_Mult.cpp:263-450) and DMEM_ChebyUpdate (src/DMEM_Misc.cpp:612-666) through
    the reference's object code on one rank (MPI of one process, oracle/ref_shim/dmem_stub.h)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import hierarchy_from_golden
    d = {}
    for name in ("lap5pt_n32", "lap7pt_n12"):
        h, g = hierarchy_from_golden(name)
        b, w = g["b"], 0.9
        for tag, sym, post in (("sym", True, 1), ("plain", False, 0)):
            h.build_transfers(H.MULTADD, w, num_pre=1, num_post=post)
            x, hist = O.ref_dmem_sync_add(h, b, w, symmetrised=sym, num_cycles=100, tol=1e-9)
            d["%s_%s_hist" % (name, tag)] = hist
            d["%s_%s_x" % (name, tag)] = x
            print(name, tag, len(hist) - 1, hist[-1])
        h.build_transfers(H.MULTADD, w)
        pb = O.Problem(h, H.MULTADD, H.JACOBI, w, coarse_solve=1)
        lo, hi = pb.eigs_power(20)
        mu, delta = (hi + lo) / (hi - lo), 2.0 / (hi + lo)
        d[name + "_mu_delta"] = np.asarray([mu, delta])
        # in the reference's (stale) src/Main.hpp:79-81 CHEBY_ACCEL == RICHARD_ACCEL == 1, so accel_type 1 takes the Richardson
        # branch of DMEM_ChebyUpdate; any other non-zero value reaches the Chebyshev recurrence
        d[name + "_richardson_hist"] = _dmem_accel_with_reference_update(h, pb, b, mu, delta, 1)
        d[name + "_chebyshev_hist"] = _dmem_accel_with_reference_update(h, pb, b, mu, delta, 2)
        print(name, "accel", len(d[name + "_richardson_hist"]) - 1, len(d[name + "_chebyshev_hist"]) - 1)
        # AddCycle + DMEM_AddSmooth, grid after grid on one rank (DMEM_Add's asynchronous loop without overlap)
        x, hist = O.ref_dmem_add_cycles(h, b, w, symmetrised=True, rounds=12)
        d[name + "_addcycle_hist"] = hist
        d[name + "_addcycle_x"] = x
        # DMEM_AsyncSmooth on one rank: 17 relaxations of the fine system, weighted and L1 Jacobi
        l1 = h.l1_norms()[0]
        d[name + "_asyncsmooth_j_x"] = O.ref_dmem_async_smooth(h.A[0], b, w, 17)[0]
        d[name + "_asyncsmooth_l1_x"] = O.ref_dmem_async_smooth(h.A[0], b, w, 17, l1=l1)[0]
    np.savez_compressed(os.path.join(OUT, "dmem.npz"), **d)


def cheby_setup_golden():
    """ChebySetup -> EigsPower -> BPXCycle (src/SMEM_Cheby.cpp:28-60,410-518,520-645) through the reference's object code
    (SMEM_Cheby.cpp compiled unmodified against SLEPc / LOBPCG stand-ins that are never reached): the eigenvalue bounds of
    B A and the Chebyshev scalars mu, delta for weighted Jacobi, L1 Jacobi and the hybrid smoother (4 threads = 4 blocks)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import hierarchy_from_golden
    d = {}
    for name in ("lap5pt_n32", "lap7pt_n12"):
        h, g = hierarchy_from_golden(name)
        h.build_transfers(H.BPX, 0.8)
        for tag, sm, w, nt in (("j", H.JACOBI, 0.8, 4), ("l1", H.L1_JACOBI, 0.8, 4), ("hjgs", H.HYBRID_JACOBI_GAUSS_SEIDEL, 1.0, 4)):
            for iters in (3, 20):
                r = O.ref_cheby_setup(h, g["b"], sm, w, iters, 1, nt)
                assert np.array_equal(r["f_after"], g["b"])
                d["%s_%s_it%d" % (name, tag, iters)] = np.asarray([r["alpha"], r["beta"], r["mu"], r["delta"]])
                print(name, tag, iters, d["%s_%s_it%d" % (name, tag, iters)])
    np.savez_compressed(os.path.join(OUT, "cheby_setup.npz"), **d)


ASYNC_CASES = (("multadd", H.ASYNC_MULTADD, H.MULTADD, 0.9, 1), ("afacx", H.ASYNC_AFACX, H.AFACX, 0.6, 1), ("afacx2", H.ASYNC_AFACX, H.AFACX, 0.6, 2))


def async_golden():
    """SMEM_Async_Add_AMG (src/SMEM_Async_AMG.cpp:7-437) through the reference's object code on the first TWO levels of the
    committed hierarchies, one thread per level: a single working group, so the asynchronous iteration is deterministic
    (the coarsest group adds exactly zero).  Multadd (symmetrised Jacobi) and AFACx (1 and 2 sweeps)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import hierarchy_from_golden
    d = {}
    for name in ("lap5pt_n32", "lap7pt_n12"):
        hf, g = hierarchy_from_golden(name)
        h = H.Hierarchy(hf.A[:2], hf.P_plain[:1])
        for tag, solver, base, w, sweeps in ASYNC_CASES:
            h.build_transfers(base, w)
            for K in (1, 7, 30):
                rs = O.RefSolver(h, solver, H.JACOBI, g["b"], w, one_thread_per_level=True, fine_sweeps=sweeps, coarse_sweeps=sweeps)
                out = rs.solve(K, 1e-9, async_type=0)
                rs.close()
                assert list(out["corrections"]) == [K, K]
                d["%s_%s_k%d_u" % (name, tag, K)] = out["u"]
                d["%s_%s_k%d_relres" % (name, tag, K)] = np.asarray(out["relres"])
                print(name, tag, K, out["relres"])
    np.savez_compressed(os.path.join(OUT, "async_two_level.npz"), **d)


def par_bpx_golden():
    """`-solver par_bpx` (PAR_BPX branch of SMEM_Sync_Parfor_BPXcycle, src/SMEM_Sync_AMG.cpp:183-236) through SMEM_Solve of the
    reference's object code, 4 threads, 12 cycles, weight 0.6: weighted Jacobi (the weight ends up squared) and L1 (step w / l1)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import hierarchy_from_golden
    d = {}
    for name in ("lap5pt_n32", "lap7pt_n12"):
        h, g = hierarchy_from_golden(name)
        h.build_transfers(H.BPX, 0.6)
        for tag, sm in (("j", H.JACOBI), ("l1", H.L1_JACOBI)):
            rs = O.RefSolver(h, H.PAR_BPX, sm, g["b"], 0.6, num_threads=4)
            out = rs.solve(12, 1e-30, async_type=0)
            rs.close()
            d["%s_%s_hist" % (name, tag)] = out["hist"]
            d["%s_%s_u" % (name, tag)] = out["u"]
            print(name, tag, out["hist"][-1])
    np.savez_compressed(os.path.join(OUT, "par_bpx.npz"), **d)


def dmem_mult_golden():
    """DMEM_Mult / DMEM_MultCycle (src/DMEM_Mult.cpp:13-261; the DMEM driver's multiplicative comparator: V(1,1), weighted Jacobi,
    direct solve on the coarsest level) through the reference's object code on one rank"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import hierarchy_from_golden
    d = {}
    for name in ("lap5pt_n32", "lap7pt_n12"):
        h, g = hierarchy_from_golden(name)
        h.build_transfers(H.MULT, 0.8)
        x, hist = O.ref_dmem_mult(h, g["b"], 0.8, 100, 1e-9)
        d[name + "_hist"], d[name + "_x"] = hist, x
        print(name, len(hist) - 1, hist[-1])
    np.savez_compressed(os.path.join(OUT, "dmem_mult.npz"), **d)


def smooth_transfer_golden():
    """SmoothTransfer (src/SMEM_Setup.cpp:1173-1254; EigenMatMat, CSR_Transpose, StdVector_to_CSR) through the reference's object
    code (SMEM_Setup.cpp compiled unmodified against the Eigen stand-in oracle/ref_shim/eigen_stub) on every level of the
    committed hierarchies: Pbar = G P, Rbar = P^T GT with the Jacobi (w = 0.9) and the L1 smoothing factors"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import hierarchy_from_golden
    d = {}
    for name in ("lap5pt_n32", "lap7pt_n12"):
        h, g = hierarchy_from_golden(name)
        for tag, kind in (("j", H.JACOBI), ("l1", H.L1_JACOBI)):
            for l in range(h.num_levels - 1):
                Pb, Rb = O.ref_smooth_transfer(h.A[l], h.P_plain[l], 0.9, kind, 1, 1)
                for mn, m in (("P", Pb), ("R", Rb)):
                    k = "%s_%s_%s%d_" % (name, tag, mn, l)
                    kp = "%s_%s%d_" % (name, mn, l)                         # the pattern does not depend on the smoothing factor
                    d[kp + "shape"] = np.asarray([m.nrows, m.ncols])
                    d[kp + "indptr"], d[kp + "indices"] = m.indptr.astype(np.int32), m.indices.astype(np.int16 if m.ncols < 32768 else np.int32)
                    if m.nnz > 2000:  # the big levels: keep the pattern and the row / column sums of the values (fixture size)
                        S = m.to_scipy()
                        d[k + "rowsum"] = np.asarray(S.sum(axis=1)).ravel()
                        d[k + "colsum"] = np.asarray(S.sum(axis=0)).ravel()
                    else:
                        d[k + "data"] = m.data
        print(name, h.num_levels)
    np.savez_compressed(os.path.join(OUT, "smooth_transfer.npz"), **d)


if __name__ == "__main__":
    if "--smooth-transfer-only" in sys.argv:
        from oracle import build as obuild
        amg.build.build_host(); obuild.build_oracle(); obuild.build_ref()
        smooth_transfer_golden()
        sys.exit(0)
    if "--dmem-mult-only" in sys.argv:
        from oracle import build as obuild
        amg.build.build_host(); obuild.build_oracle(); obuild.build_ref()
        dmem_mult_golden()
        sys.exit(0)
    if "--par-bpx-only" in sys.argv:
        from oracle import build as obuild
        amg.build.build_host(); obuild.build_oracle(); obuild.build_ref()
        par_bpx_golden()
        sys.exit(0)
    if "--async-only" in sys.argv:
        from oracle import build as obuild
        amg.build.build_host(); obuild.build_oracle(); obuild.build_ref()
        async_golden()
        sys.exit(0)
    if "--cheby-setup-only" in sys.argv:
        from oracle import build as obuild
        amg.build.build_host(); obuild.build_oracle(); obuild.build_ref()
        cheby_setup_golden()
        sys.exit(0)
    if "--dmem-only" in sys.argv:
        from oracle import build as obuild
        amg.build.build_host(); obuild.build_oracle(); obuild.build_ref()
        dmem_golden()
        sys.exit(0)
    if "--cheby-only" in sys.argv:
        from oracle import build as obuild
        amg.build.build_host(); obuild.build_oracle(); obuild.build_ref()
        cheby_golden()
        sys.exit(0)
    if "--hybrid-jgs-only" in sys.argv:
        from oracle import build as obuild
        amg.build.build_host(); obuild.build_oracle(); obuild.build_ref()
        hybrid_jgs_golden()
        sys.exit(0)
    if "--iebpx-only" in sys.argv:
        from oracle import build as obuild
        amg.build.build_host(); obuild.build_oracle(); obuild.build_ref()
        iebpx_golden()
        sys.exit(0)
    if "--matrix-file-only" not in sys.argv:
        main()
    else:
        from oracle import build as obuild
        amg.build.build_host(); obuild.build_ref()
    matrix_file_golden()
    if "--matrix-file-only" not in sys.argv:
        iebpx_golden()
        hybrid_jgs_golden()
        cheby_golden()
        cheby_setup_golden()
        smooth_transfer_golden()
        par_bpx_golden()
        dmem_mult_golden()
        async_golden()
        dmem_golden()
