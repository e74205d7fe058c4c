"""ctypes front-end of the CPU oracle (oracle/amg_oracle.c) and of the compiled reference
objects (oracle/_ref/libref_smem.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
import ctypes as C
import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("async-multigrid_b200")
from oracle import build as _build  # noqa: E402

DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)


class OrcCSR(C.Structure):
    _fields_ = [("nrows", C.c_int), ("ncols", C.c_int), ("nnz", C.c_int),
                ("i", IP), ("j", IP), ("data", DP)]


class OrcProblem(C.Structure):
    _fields_ = [("num_levels", C.c_int),
                ("A", C.POINTER(OrcCSR)), ("P", C.POINTER(OrcCSR)), ("R", C.POINTER(OrcCSR)),
                ("l1", C.POINTER(DP)),
                ("solver", C.c_int), ("smoother", C.c_int), ("smooth_weight", C.c_double),
                ("num_pre", C.c_int), ("num_post", C.c_int), ("fine_sweeps", C.c_int), ("coarse_sweeps", C.c_int),
                ("jgs_blocks", C.POINTER(IP)), ("jgs_nblocks", IP), ("jgs_parfor_scale", C.c_int),
                ("coarse_solve", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.ORACLE_LIB
        if not os.path.exists(path):
            _build.build_oracle()
        L = C.CDLL(path)
        L.orc_norm2.restype = C.c_double
        L.orc_norm2.argtypes = [DP, C.c_int]
        L.orc_solve_sync.restype = C.c_int
        L.orc_solve_sync.argtypes = [C.POINTER(OrcProblem), DP, DP, C.c_double, C.c_int, C.c_int,
                                     C.c_double, C.c_double, DP, DP]
        L.orc_cycle.argtypes = [C.POINTER(OrcProblem), DP, DP]
        L.orc_spgemv.argtypes = [C.POINTER(OrcCSR), DP, DP, C.c_double, C.c_double, DP]
        L.orc_matvec.argtypes = [C.POINTER(OrcCSR), DP, DP, C.c_int, C.c_int]
        L.orc_jacobi.argtypes = [C.POINTER(OrcCSR), DP, DP, DP, C.c_double, C.c_int, C.c_int]
        L.orc_l1_jacobi.argtypes = [C.POINTER(OrcCSR), DP, DP, DP, DP, C.c_int, C.c_int]
        L.orc_hybrid_jgs.argtypes = [C.POINTER(OrcCSR), DP, DP, DP, IP, C.c_int, DP, C.c_int, C.c_int]
        L.orc_symmetric_jacobi.argtypes = [C.POINTER(OrcCSR), DP, DP, DP, DP, C.c_double, C.c_int, C.c_int]
        L.orc_symmetric_l1_jacobi.argtypes = [C.POINTER(OrcCSR), DP, DP, DP, DP, DP, C.c_int, C.c_int]
        L.orc_solve_async_sequential.argtypes = [C.POINTER(OrcProblem), DP, DP, C.c_int, IP, DP]
        L.orc_solve_sync_dmem.restype = C.c_int
        L.orc_solve_sync_dmem.argtypes = [C.POINTER(OrcProblem), DP, DP, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, DP]
        L.orc_eigs_power.argtypes = [C.POINTER(OrcProblem), C.c_int, DP, DP]
        L.orc_eigs_power.restype = None
        _lib = L
    return _lib


def dptr(a):
    return a.ctypes.data_as(DP)


def iptr(a):
    return a.ctypes.data_as(IP)


def c_csr(m):
    s = OrcCSR()
    s.nrows, s.ncols, s.nnz = m.nrows, m.ncols, m.nnz
    s.i, s.j, s.data = iptr(m.indptr), iptr(m.indices), dptr(m.data)
    return s


class Problem:
    """Keeps the numpy arrays alive and exposes an orc_problem for the C side."""

    def __init__(self, h, solver, smoother, smooth_weight=1.0, num_pre=1, num_post=1,
                 fine_sweeps=1, coarse_sweeps=1, jgs_blocks=None, jgs_parfor_scale=0, coarse_solve=0, l1_scale=1.0):
        L = h.num_levels
        self.h = h
        self._keep = (list(h.A), list(h.P), list(h.R))   # the C structs borrow these arrays
        self._A = (OrcCSR * L)(*[c_csr(a) for a in h.A])
        self._P = (OrcCSR * max(L - 1, 1))(*[c_csr(p) for p in h.P])
        self._R = (OrcCSR * max(L - 1, 1))(*[c_csr(r) for r in h.R])
        self._l1arrs = [np.ascontiguousarray(a * l1_scale) for a in h.l1_norms()]     # l1_scale = 1 / w: the step w / l1 of PAR_BPX
        self._l1 = (DP * L)(*[dptr(a) for a in self._l1arrs])
        if jgs_blocks is None:
            jgs_blocks = [np.asarray([0, a.nrows], dtype=np.int32) for a in h.A]
        self._blocks = [np.ascontiguousarray(b, dtype=np.int32) for b in jgs_blocks]
        self._bptr = (IP * L)(*[iptr(b) for b in self._blocks])
        self._nb = np.asarray([len(b) - 1 for b in self._blocks], dtype=np.int32)
        p = OrcProblem()
        p.num_levels = L
        p.A, p.P, p.R = self._A, self._P, self._R
        p.l1 = self._l1
        p.solver, p.smoother, p.smooth_weight = solver, smoother, smooth_weight
        p.num_pre, p.num_post, p.fine_sweeps, p.coarse_sweeps = num_pre, num_post, fine_sweeps, coarse_sweeps
        p.jgs_blocks = self._bptr
        p.jgs_nblocks = iptr(self._nb)
        p.jgs_parfor_scale = jgs_parfor_scale
        p.coarse_solve = int(coarse_solve)
        self.c = p

    def solve_sync(self, f, tol=1e-9, num_cycles=100, cheby=None, u0=None):
        n = self.h.n[0]
        u = np.zeros(n) if u0 is None else np.array(u0, dtype=np.float64)
        hist = np.zeros(num_cycles + 1)
        secs = C.c_double(0)
        mu, delta = (cheby if cheby else (0.0, 0.0))
        k = lib().orc_solve_sync(C.byref(self.c), dptr(np.ascontiguousarray(f)), dptr(u), tol, num_cycles,
                                 1 if cheby else 0, mu, delta, dptr(hist), C.byref(secs))
        return u, hist[:k + 1], secs.value

    def cycle(self, r):
        u = np.zeros(self.h.n[0])
        lib().orc_cycle(C.byref(self.c), dptr(np.ascontiguousarray(r)), dptr(u))
        return u

    def solve_sync_dmem(self, f, tol=1e-9, num_cycles=100, accel=0, mu=1.0, delta=1.0):
        """DMEM_SyncAddCorrect + DMEM_ChebyUpdate: accel 0 none, 1 Chebyshev, 2 second-order Richardson"""
        u = np.zeros(self.h.n[0])
        hist = np.zeros(num_cycles + 1)
        k = lib().orc_solve_sync_dmem(C.byref(self.c), dptr(np.ascontiguousarray(f)), dptr(u), tol, num_cycles, accel, mu, delta, dptr(hist))
        return u, hist[:k + 1]

    def eigs_power(self, iters=20):
        """(alpha, beta) = (eig_min, eig_max) of B*A as EigsPower estimates them"""
        a, b = C.c_double(0), C.c_double(0)
        lib().orc_eigs_power(C.byref(self.c), iters, C.byref(a), C.byref(b))
        return a.value, b.value

    def solve_iebpx(self, f, tol=1e-9, num_cycles=100, mu=1.0, delta=1.0):
        """implicit extended-system BPX (`-solver iebpx`, src/SMEM_ExtendedSystem.cpp) -> dict(x, iters = the reference's
        loc_iters, ext_hist[1..iters-1], ext_relres, relres)"""
        x = np.zeros(self.h.n[0])
        hist = np.zeros(max(num_cycles, 2) + 1)
        er, rr = C.c_double(0), C.c_double(0)
        lib().orc_solve_iebpx.restype = C.c_int
        it = lib().orc_solve_iebpx(C.byref(self.c), dptr(np.ascontiguousarray(f)), dptr(x), C.c_double(tol), int(num_cycles),
                                   C.c_double(mu), C.c_double(delta), dptr(hist), C.byref(er), C.byref(rr))
        return dict(x=x, iters=it, ext_hist=hist[:it], ext_relres=er.value, relres=rr.value)

    def solve_async_sequential(self, f, num_cycles, read_res=False):
        """read_res: `-read_type res` (the shared residual is updated incrementally, u is assembled at the end)"""
        u = np.zeros(self.h.n[0])
        counts = np.zeros(self.h.num_levels, dtype=np.int32)
        rr = C.c_double(0)
        fn = lib().orc_solve_async_sequential_res if read_res else lib().orc_solve_async_sequential
        fn.argtypes = [C.POINTER(OrcProblem), DP, DP, C.c_int, IP, DP]
        fn(C.byref(self.c), dptr(np.ascontiguousarray(f)), dptr(u), num_cycles, iptr(counts), C.byref(rr))
        return u, counts, rr.value


def spgemv(m, x, b, alpha, beta):
    y = np.zeros(m.nrows)
    s = c_csr(m)
    bb = b if b is not None else np.zeros(m.nrows)
    lib().orc_spgemv(C.byref(s), dptr(np.ascontiguousarray(x)), dptr(np.ascontiguousarray(bb)), alpha, beta, dptr(y))
    return y


def matvecT(m, x):
    """y = A^T x, sequential scatter (src/SEQ_MatVec.cpp:26-45)"""
    y = np.zeros(m.ncols)
    s = c_csr(m)
    lib().orc_matvecT.argtypes = [C.POINTER(OrcCSR), DP, DP]
    lib().orc_matvecT(C.byref(s), dptr(np.ascontiguousarray(x, dtype=np.float64)), dptr(y))
    return y


def ref_parfor_matvec_t(m, x, num_threads=4):
    """SMEM_Sync_Parfor_MatVecT (src/SMEM_MatVec.cpp:27-58), the reference's object code"""
    L = ref_lib()
    y = np.zeros(m.ncols)
    s = c_csr(m)
    L.ref_parfor_matvec_t.restype = None
    L.ref_parfor_matvec_t.argtypes = [C.POINTER(OrcCSR), DP, DP, C.c_int]
    L.ref_parfor_matvec_t(C.byref(s), dptr(np.ascontiguousarray(x, dtype=np.float64)), dptr(y), num_threads)
    return y


def norm2(x):
    x = np.ascontiguousarray(x)
    return lib().orc_norm2(dptr(x), len(x))


def smooth(kind, m, f, w=1.0, sweeps=1, zero_flag=1, u0=None, l1=None, blocks=None, scale=None):
    """kind in jacobi | l1_jacobi | hybrid_jgs | symmetric_jacobi | symmetric_l1_jacobi"""
    n = m.nrows
    s = c_csr(m)
    f = np.ascontiguousarray(f)
    u = np.zeros(n) if u0 is None else np.array(u0, dtype=np.float64)
    y, r = np.zeros(n), np.zeros(n)
    L = lib()
    if kind == "jacobi":
        L.orc_jacobi(C.byref(s), dptr(f), dptr(u), dptr(y), w, sweeps, zero_flag)
    elif kind == "l1_jacobi":
        L.orc_l1_jacobi(C.byref(s), dptr(f), dptr(u), dptr(y), dptr(l1), sweeps, zero_flag)
    elif kind == "hybrid_jgs":
        blocks = np.ascontiguousarray(blocks, dtype=np.int32)
        sp = dptr(scale) if scale is not None else None
        L.orc_hybrid_jgs(C.byref(s), dptr(f), dptr(u), dptr(y), iptr(blocks), len(blocks) - 1, sp, sweeps, zero_flag)
    elif kind == "symmetric_jacobi":
        L.orc_symmetric_jacobi(C.byref(s), dptr(f), dptr(u), dptr(y), dptr(r), w, sweeps, zero_flag)
    elif kind == "symmetric_l1_jacobi":
        L.orc_symmetric_l1_jacobi(C.byref(s), dptr(f), dptr(u), dptr(y), dptr(r), dptr(l1), sweeps, zero_flag)
    else:
        raise ValueError(kind)
    return u


# ------------------------------------------------------------------------------------------------
# the reference's own object code (oracle/_ref/libref_smem.so, built by oracle/build_ref.sh)
# ------------------------------------------------------------------------------------------------
_ref = None


_ref_b200 = None


def ref_b200_lib():
    """oracle/_ref/libref_b200.so: the reference's object code + the binding of INTEGRATION.md (integration/SMEM_B200.hpp),
    linked against libamg_b200.so -- the reference-side structs drive the product library.  None when absent; loading it
    needs the CUDA runtime (GPU box)."""
    global _ref_b200
    if _ref_b200 is None:
        path = _build.REF_B200_LIB
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        _declare_ref(L)
        L.ref_solve_b200.restype = C.c_int
        L.ref_solve_b200.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, DP, DP, IP, DP]
        _ref_b200 = L
    return _ref_b200


def ref_lib():
    """None when oracle/_ref is absent (it is built only where /root/reference is mounted and
    travels to the GPU box as a prebuilt file)."""
    global _ref
    if _ref is None:
        path = _build.REF_LIB
        if not os.path.exists(path):
            try:
                _build.build_ref()
            except Exception:
                return None
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        _declare_ref(L)
        _ref = L
    return _ref


def _declare_ref(L):
    if True:
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.POINTER(DP),
                                 C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, IP, DP]
        L.ref_solve.restype = C.c_int
        L.ref_solve.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                C.c_int, DP, DP, IP, DP, DP]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_solve_sync_det.restype = C.c_int
        L.ref_solve_sync_det.argtypes = [C.c_void_p, C.c_int, C.c_double, DP, DP, DP]
        L.ref_matvec.argtypes = [C.POINTER(OrcCSR), DP, DP]
        L.ref_seq_symmetric_jacobi.argtypes = [C.POINTER(OrcCSR), DP, DP, C.c_double, C.c_int]
        L.ref_seq_jacobi.argtypes = [C.POINTER(OrcCSR), DP, DP, C.c_double, C.c_int, C.c_int]
        L.ref_read_matrix.restype = C.c_int
        L.ref_read_matrix.argtypes = [C.c_char_p, C.c_int, IP, IP, C.POINTER(IP), C.POINTER(IP), C.POINTER(DP)]
        L.ref_free.argtypes = [C.c_void_p]


def ref_solve_eebpx(h, AA, disp, bb, num_cycles, tol=1e-9, mu=1.0, delta=1.0, num_threads=4):
    """the reference's SMEM_ExtendedSystemSolve, EXPLICIT_EXTENDED_SYSTEM_BPX branch (object code in oracle/_ref), on an
    assembled extended matrix -> dict(x, xx, iters, ext_relres, relres)"""
    L = ref_lib()
    nl = h.num_levels
    keep = (list(h.A), list(h.P), AA)
    A = (OrcCSR * nl)(*[c_csr(a) for a in h.A])
    P = (OrcCSR * max(nl - 1, 1))(*[c_csr(p) for p in h.P])
    aa = c_csr(AA)
    x, xx = np.zeros(h.n[0]), np.zeros(AA.nrows)
    er, rr = C.c_double(0), C.c_double(0)
    L.ref_solve_eebpx.restype = C.c_int
    L.ref_solve_eebpx.argtypes = [C.c_int, C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.POINTER(OrcCSR), IP, DP, C.c_int, C.c_int,
                                  C.c_double, C.c_double, C.c_double, DP, DP, DP, DP]
    d = np.ascontiguousarray(disp, dtype=np.int32)
    it = L.ref_solve_eebpx(nl, A, P, C.byref(aa), iptr(d), dptr(np.ascontiguousarray(bb, dtype=np.float64)), num_threads,
                           num_cycles, tol, mu, delta, dptr(x), dptr(xx), C.byref(er), C.byref(rr))
    del keep
    return dict(x=x, xx=xx, iters=it, ext_relres=er.value, relres=rr.value)


def ref_dmem_sync_add(h, b, smooth_weight, symmetrised=True, num_cycles=100, tol=1e-9):
    """the reference's DMEM_SyncAdd / DMEM_SyncAddCycle object code (src/DMEM_Mult.cpp:263-450) on one rank, Multadd with the
    DMEM conventions (direct solve on the coarsest level); h.P / h.R are the smoothed transfers.  -> (x, hist)"""
    L = ref_lib()
    nl = h.num_levels
    Rt = [_pkg.hierarchy.CSR.from_scipy(r.to_scipy().T.tocsr()) for r in h.R]     # hypre applies R_array transposed
    keep = (list(h.A), list(h.P), Rt)
    A = (OrcCSR * nl)(*[c_csr(a) for a in h.A])
    P = (OrcCSR * max(nl - 1, 1))(*[c_csr(p) for p in h.P])
    R = (OrcCSR * max(nl - 1, 1))(*[c_csr(r) for r in Rt])
    x, hist = np.zeros(h.n[0]), np.zeros(num_cycles + 1)
    L.ref_dmem_sync_add.restype = C.c_int
    L.ref_dmem_sync_add.argtypes = [C.c_int, C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.c_double, C.c_int, DP,
                                    C.c_int, C.c_double, DP, DP]
    k = L.ref_dmem_sync_add(nl, A, P, R, smooth_weight, int(symmetrised), dptr(np.ascontiguousarray(b, dtype=np.float64)),
                            num_cycles, tol, dptr(x), dptr(hist))
    del keep
    return x, hist[:k + 1]


def ref_dmem_solve_b200(h, b, smooth_weight, symmetrised=True, num_cycles=100, tol=1e-9, async_flag=0):
    """the DMEM binding of INTEGRATION.md (integration/DMEM_B200.hpp compiled against the reference's DMEM_Main.hpp) on one rank:
    DMEM_AllData -> DMEM_B200_Upload -> DMEM_Add_B200.  h.P / h.R are the smoothed transfers.  Needs the GPU (libref_b200.so is
    linked against the product library).  -> dict(x, hist, cycles, corrections, relres)"""
    L = ref_b200_lib()
    if L is None:
        return None
    nl = h.num_levels
    keep = (list(h.A), list(h.P), list(h.R))
    A = (OrcCSR * nl)(*[c_csr(a) for a in h.A])
    P = (OrcCSR * max(nl - 1, 1))(*[c_csr(p) for p in h.P])
    R = (OrcCSR * max(nl - 1, 1))(*[c_csr(r) for r in h.R])
    x, hist = np.zeros(h.n[0]), np.zeros(num_cycles + 1)
    cor = np.zeros(nl, dtype=np.int32)
    rel = C.c_double(0)
    L.ref_dmem_solve_b200.restype = C.c_int
    L.ref_dmem_solve_b200.argtypes = [C.c_int, C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.c_double, C.c_int, DP, C.c_int,
                                      C.c_double, C.c_int, DP, DP, IP, DP]
    k = L.ref_dmem_solve_b200(nl, A, P, R, smooth_weight, int(symmetrised), dptr(np.ascontiguousarray(b, dtype=np.float64)), num_cycles,
                              tol, int(async_flag), dptr(x), dptr(hist), iptr(cor), C.byref(rel))
    del keep
    return dict(x=x, hist=hist[:k + 1] if not async_flag else None, cycles=k, corrections=cor, relres=rel.value)


def ref_dmem_mult(h, b, smooth_weight, num_cycles=100, tol=1e-9):
    """the reference's DMEM_Mult / DMEM_MultCycle object code (src/DMEM_Mult.cpp:13-261) on one rank: multiplicative V(1,1),
    weighted Jacobi, direct solve on the coarsest level; h.P / h.R plain.  -> (x, hist)"""
    L = ref_lib()
    nl = h.num_levels
    Rt = [_pkg.hierarchy.CSR.from_scipy(r.to_scipy().T.tocsr()) for r in h.R]     # hypre applies R_array transposed
    keep = (list(h.A), list(h.P), Rt)
    A = (OrcCSR * nl)(*[c_csr(a) for a in h.A])
    P = (OrcCSR * max(nl - 1, 1))(*[c_csr(p) for p in h.P])
    R = (OrcCSR * max(nl - 1, 1))(*[c_csr(r) for r in Rt])
    x, hist = np.zeros(h.n[0]), np.zeros(num_cycles + 1)
    L.ref_dmem_mult.restype = C.c_int
    L.ref_dmem_mult.argtypes = [C.c_int, C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.c_double, DP, C.c_int, C.c_double,
                                DP, DP]
    k = L.ref_dmem_mult(nl, A, P, R, smooth_weight, dptr(np.ascontiguousarray(b, dtype=np.float64)), num_cycles, tol, dptr(x),
                        dptr(hist))
    del keep
    return x, hist[:k + 1]


def ref_dmem_add_cycles(h, b, smooth_weight, symmetrised=True, rounds=10):
    """AddCycle + DMEM_AddSmooth (src/DMEM_Add.cpp:180-329, src/DMEM_Smooth.cpp:574-638), the reference's object code, grid
    after grid on one rank -> (x, hist per round)"""
    L = ref_lib()
    nl = h.num_levels
    Rt = [_pkg.hierarchy.CSR.from_scipy(r.to_scipy().T.tocsr()) for r in h.R]
    keep = (list(h.A), list(h.P), Rt)
    A = (OrcCSR * nl)(*[c_csr(a) for a in h.A])
    P = (OrcCSR * max(nl - 1, 1))(*[c_csr(p) for p in h.P])
    R = (OrcCSR * max(nl - 1, 1))(*[c_csr(r) for r in Rt])
    x, hist = np.zeros(h.n[0]), np.zeros(rounds + 1)
    L.ref_dmem_add_cycles.restype = C.c_int
    L.ref_dmem_add_cycles.argtypes = [C.c_int, C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.c_double, C.c_int, DP,
                                      C.c_int, DP, DP]
    L.ref_dmem_add_cycles(nl, A, P, R, smooth_weight, int(symmetrised), dptr(np.ascontiguousarray(b, dtype=np.float64)), rounds,
                          dptr(x), dptr(hist))
    del keep
    return x, hist


def ref_dmem_async_smooth(A, b, smooth_weight, num_cycles, l1=None):
    """DMEM_AsyncSmooth (src/DMEM_Smooth.cpp:16-313), the reference's object code, on one rank (no neighbour): ASYNC_JACOBI, or
    ASYNC_L1_JACOBI when l1 is given -> (x, r as the reference maintains it incrementally, relaxations done)"""
    L = ref_lib()
    a = c_csr(A)
    x, r = np.zeros(A.nrows), np.zeros(A.nrows)
    l1a = np.ascontiguousarray(l1 if l1 is not None else np.ones(A.nrows), dtype=np.float64)
    L.ref_dmem_async_smooth.restype = C.c_int
    L.ref_dmem_async_smooth.argtypes = [C.POINTER(OrcCSR), DP, C.c_double, DP, C.c_int, C.c_int, DP, DP]
    k = L.ref_dmem_async_smooth(C.byref(a), dptr(np.ascontiguousarray(b, dtype=np.float64)), smooth_weight, dptr(l1a),
                                10 if l1 is not None else 8, num_cycles, dptr(x), dptr(r))
    return x, r, k


def ref_cheby_setup(h, b, smoother, smooth_weight, iters=20, num_sweeps=1, num_threads=1):
    """ChebySetup -> EigsPower -> BPXCycle (src/SMEM_Cheby.cpp:28-60,410-518,520-645), the reference's object code, on the plain
    interpolants h.P (hypre's R_array is P_array, applied transposed) -> dict(alpha, beta, mu, delta, f_after)"""
    L = ref_lib()
    nl = h.num_levels
    l1arrs = h.l1_norms()
    keep = (list(h.A), list(h.P), l1arrs)
    A = (OrcCSR * nl)(*[c_csr(a) for a in h.A])
    P = (OrcCSR * max(nl - 1, 1))(*[c_csr(p) for p in h.P])
    l1 = (DP * nl)(*[dptr(a) for a in l1arrs])
    out, fa = np.zeros(4), np.zeros(h.n[0])
    L.ref_cheby_setup.restype = C.c_int
    L.ref_cheby_setup.argtypes = [C.c_int, C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.POINTER(DP), C.c_int, C.c_double, C.c_int,
                                  C.c_int, C.c_int, DP, DP, DP]
    rc = L.ref_cheby_setup(nl, A, P, l1, smoother, smooth_weight, num_sweeps, iters, num_threads,
                           dptr(np.ascontiguousarray(b, dtype=np.float64)), dptr(out), dptr(fa))
    del keep
    assert rc == 0
    return dict(alpha=out[0], beta=out[1], mu=out[2], delta=out[3], f_after=fa)


def ref_smooth_transfer(A, P, smooth_weight, smooth_interp_type=0, num_pre=1, num_post=1, num_threads=2):
    """SmoothTransfer (src/SMEM_Setup.cpp:1173-1254, + EigenMatMat / CSR_Transpose / StdVector_to_CSR), the reference's object code
    compiled against an Eigen stand-in, for one level -> (Pbar or None, Rbar or None) as hierarchy.CSR"""
    L = ref_lib()
    hier = _pkg.hierarchy
    a, p = c_csr(A), c_csr(P)
    l1 = np.ascontiguousarray(np.add.reduceat(np.abs(A.data), A.indptr[:-1]) if A.nnz else np.zeros(A.nrows))
    oP, oR = OrcCSR(), OrcCSR()
    L.ref_smooth_transfer.restype = C.c_int
    L.ref_smooth_transfer.argtypes = [C.POINTER(OrcCSR), C.POINTER(OrcCSR), DP, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(OrcCSR), C.POINTER(OrcCSR)]
    L.ref_smooth_transfer(C.byref(a), C.byref(p), dptr(l1), smooth_weight, smooth_interp_type, num_pre, num_post, num_threads,
                          C.byref(oP), C.byref(oR))
    L.ref_free.argtypes = [C.c_void_p]

    def take(o):
        if o.nrows < 0:
            return None
        n, nnz = o.nrows, o.nnz
        ip = np.ctypeslib.as_array(o.i, shape=(n + 1,)).copy()
        ix = np.ctypeslib.as_array(o.j, shape=(nnz,)).copy()
        va = np.ctypeslib.as_array(o.data, shape=(nnz,)).copy()
        for ptr in (o.i, o.j, o.data):
            L.ref_free(C.cast(ptr, C.c_void_p))
        return hier.CSR(n, o.ncols, ip, ix, va)
    return take(oP), take(oR)


def ref_work_partition(h, solver, num_threads, num_pre=1, num_post=1, fine_sweeps=1, coarse_sweeps=1):
    """ComputeWork + PartitionLevels (BALANCED_THREADS) + PartitionGrids of the reference's object code (src/SMEM_Setup.cpp:590-1170)
    -> dict(level_work, frac, threads_per_level, A_ns, A_ne [L x num_threads, -1 = not set])"""
    L = ref_lib()
    nl = h.num_levels
    keep = (list(h.A), list(h.P), list(h.R))
    A = (OrcCSR * nl)(*[c_csr(a) for a in h.A])
    P = (OrcCSR * max(nl - 1, 1))(*[c_csr(p) for p in h.P])
    R = (OrcCSR * max(nl - 1, 1))(*[c_csr(r) for r in h.R])
    lw, fr, tpl = np.zeros(nl, dtype=np.int32), np.zeros(nl), np.zeros(nl, dtype=np.int32)
    ns, ne = np.zeros(nl * num_threads, dtype=np.int32), np.zeros(nl * num_threads, dtype=np.int32)
    L.ref_work_partition.restype = C.c_int
    L.ref_work_partition.argtypes = [C.c_int, C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, IP, DP, IP, IP, IP]
    L.ref_work_partition(nl, A, P, R, solver, num_pre, num_post, fine_sweeps, coarse_sweeps, num_threads, iptr(lw), dptr(fr), iptr(tpl),
                         iptr(ns), iptr(ne))
    del keep
    return dict(level_work=lw, frac=fr, threads_per_level=tpl, A_ns=ns.reshape(nl, num_threads), A_ne=ne.reshape(nl, num_threads))


def ref_build_extended_matrix(h):
    """BuildExtendedMatrix (src/SMEM_Setup.cpp:1426-1521), EXPLICIT_EXTENDED_SYSTEM_BPX, the reference's object code on h.A / h.P /
    h.R (plain transfers) -> (AA as hierarchy.CSR, disp)"""
    L = ref_lib()
    nl = h.num_levels
    keep = (list(h.A), list(h.P), list(h.R))
    A = (OrcCSR * nl)(*[c_csr(a) for a in h.A])
    P = (OrcCSR * max(nl - 1, 1))(*[c_csr(p) for p in h.P])
    R = (OrcCSR * max(nl - 1, 1))(*[c_csr(r) for r in h.R])
    o = OrcCSR()
    disp = np.zeros(nl + 1, dtype=np.int32)
    L.ref_build_extended_matrix.restype = C.c_int
    L.ref_build_extended_matrix.argtypes = [C.c_int, C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.POINTER(OrcCSR), C.POINTER(OrcCSR), IP]
    L.ref_build_extended_matrix(nl, A, P, R, C.byref(o), iptr(disp))
    del keep
    L.ref_free.argtypes = [C.c_void_p]
    n, nnz = o.nrows, o.nnz
    ip = np.ctypeslib.as_array(o.i, shape=(n + 1,)).copy()
    ix = np.ctypeslib.as_array(o.j, shape=(nnz,)).copy()
    va = np.ctypeslib.as_array(o.data, shape=(nnz,)).copy()
    for ptr in (o.i, o.j, o.data):
        L.ref_free(C.cast(ptr, C.c_void_p))
    return _pkg.hierarchy.CSR(n, o.ncols, ip, ix, va), disp


def ref_stencil_values(test_problem, n, c=(1.0, 1.0, 1.0), a=(0.0, 0.0, 0.0), atype=0):
    """the stencil coefficients src/BuildHypreMatrix.cpp:100-289 (the reference's object code) hands to hypre's generators:
    test_problem 1 = 7pt -> [centre, x, y, z]; 2 = 27pt -> [centre, neighbour]; 7 = difconv -> [centre, x-1, y-1, z-1, x+1, y+1, z+1]"""
    L = ref_lib()
    out = np.zeros(8)
    L.ref_stencil_values.restype = C.c_int
    L.ref_stencil_values.argtypes = [C.c_int] * 4 + [C.c_double] * 6 + [C.c_int, DP]
    k = L.ref_stencil_values(test_problem, n, n, n, c[0], c[1], c[2], a[0], a[1], a[2], atype, dptr(out))
    return out[:k]


def ref_dmem_cheby_update(d, u, cycle, mu, delta, c, c_prev, accel_type=1):
    """DMEM_ChebyUpdate (src/DMEM_Misc.cpp:612-666), synchronous branch, in place on copies -> (d, u, c, c_prev)"""
    L = ref_lib()
    d, u = np.array(d, dtype=np.float64), np.array(u, dtype=np.float64)
    cc, cp = C.c_double(c), C.c_double(c_prev)
    L.ref_dmem_cheby_update.restype = None
    L.ref_dmem_cheby_update.argtypes = [C.c_int, DP, DP, C.c_int, C.c_double, C.c_double, C.c_int, DP, DP]
    L.ref_dmem_cheby_update(len(d), dptr(d), dptr(u), cycle, mu, delta, accel_type, C.byref(cc), C.byref(cp))
    return d, u, cc.value, cp.value


def ref_read_matrix(path, symm_flag=1):
    """the reference's own reader (ReadBinary_fread_HypreParCSR, src/Misc.cpp:800-915) -> (indptr, indices, data)"""
    L = ref_lib()
    n, nnz = C.c_int(), C.c_int()
    ri, rj, rd = IP(), IP(), DP()
    if L.ref_read_matrix(os.fsencode(path), symm_flag, C.byref(n), C.byref(nnz), C.byref(ri), C.byref(rj), C.byref(rd)) != 0:
        raise IOError(path)
    ip = np.ctypeslib.as_array(ri, shape=(n.value + 1,)).copy()
    ix = np.ctypeslib.as_array(rj, shape=(max(nnz.value, 1),))[:nnz.value].copy()
    dv = np.ctypeslib.as_array(rd, shape=(max(nnz.value, 1),))[:nnz.value].copy()
    for q in (ri, rj, rd):
        L.ref_free(C.cast(q, C.c_void_p))
    return ip, ix, dv


class RefSolver:
    """SMEM_Solve of the reference on a given hierarchy.  num_threads >= num_levels is required
    for the thread-group-per-level cycles."""

    def __init__(self, h, solver, smoother, f, smooth_weight=1.0, num_pre=1, num_post=1,
                 fine_sweeps=1, coarse_sweeps=1, num_threads=None, one_thread_per_level=False, lib=None):
        hier = _pkg.hierarchy
        L = h.num_levels
        self.h = h
        self.L = ref_lib() if lib is None else lib
        if self.L is None:
            raise RuntimeError("oracle/_ref/libref_smem.so not available")
        if num_threads is None:
            num_threads = max(L, min(self.L.ref_max_threads(), 64))
        self.num_threads = num_threads
        self._keep = (list(h.A), list(h.P), list(h.R))   # the C structs borrow these arrays
        self._A = (OrcCSR * L)(*[c_csr(a) for a in h.A])
        self._P = (OrcCSR * max(L - 1, 1))(*[c_csr(p) for p in h.P])
        self._R = (OrcCSR * max(L - 1, 1))(*[c_csr(r) for r in h.R])
        self._l1arrs = h.l1_norms()
        self._l1 = (DP * L)(*[dptr(a) for a in self._l1arrs])
        _, frac = hier.compute_work(h, solver, num_pre, num_post, fine_sweeps, coarse_sweeps)
        if one_thread_per_level:
            num_threads = L
            self.threads_per_level = np.ones(L, dtype=np.int32)
        else:
            tpl = hier.balanced_threads(frac, num_threads)
            # the reference's balancing loop can leave a level with 0 threads, which is then silently never
            # corrected (the `max_diff < 0.0 && threads == 1` guard at
            # src/SMEM_Setup.cpp:788 tests the running maximum, not the candidate).  A converging run needs
            # every level served: move one thread from the best-provisioned level to each empty one.
            while min(tpl) == 0 and num_threads >= L:    # (ONE_LEVEL solvers -- MULT, BPX -- ignore the assignment)
                tpl[int(np.argmax(tpl))] -= 1
                tpl[tpl.index(0)] += 1
            self.threads_per_level = np.asarray(tpl, dtype=np.int32)
        self.num_threads = num_threads
        self._f = np.ascontiguousarray(f, dtype=np.float64)
        self.handle = self.L.ref_create(L, self._A, self._P, self._R, self._l1, solver, smoother, smooth_weight,
                                        num_pre, num_post, fine_sweeps, coarse_sweeps, num_threads,
                                        iptr(self.threads_per_level), dptr(self._f))

    def solve(self, num_cycles, tol=1e-9, async_type=1, cheby=None, precond=0):
        """async_type=1 (SEMI_ASYNC) takes the lock around the u update in the sync additive cycle
        (src/SMEM_Sync_AMG.cpp:598-610), i.e. the race-free meaning."""
        n = self.h.n[0]
        u = np.zeros(n)
        hist = np.zeros(num_cycles + 1)
        corr = np.zeros(self.h.num_levels, dtype=np.int32)
        secs, rr = C.c_double(0), C.c_double(0)
        mu, delta = cheby if cheby else (1.0, 1.0)
        k = self.L.ref_solve(self.handle, num_cycles, tol, async_type, 1 if cheby else 0, mu, delta, precond,
                             dptr(u), dptr(hist), iptr(corr), C.byref(secs), C.byref(rr))
        return dict(u=u, hist=hist[:k + 1], cycles=k, corrections=corr, seconds=secs.value, relres=rr.value)

    def solve_b200(self, num_cycles, tol=1e-9, async_type=0, res_compute_type=0, read_type=0):
        """(lib = ref_b200_lib()) InitSolve + INTEGRATION.md's SMEM_B200_Upload / SMEM_Solve_B200 on this handle's AllData"""
        n = self.h.n[0]
        u = np.zeros(n)
        hist = np.zeros(num_cycles + 1)
        corr = np.zeros(self.h.num_levels, dtype=np.int32)
        rr = C.c_double(0)
        k = self.L.ref_solve_b200(self.handle, num_cycles, tol, async_type, res_compute_type, read_type, dptr(u), dptr(hist),
                                  iptr(corr), C.byref(rr))
        return dict(u=u, hist=hist[:k + 1], cycles=k, corrections=corr, relres=rr.value)

    def solve_iebpx(self, num_cycles, tol=1e-9, mu=1.0, delta=1.0):
        """SMEM_ExtendedSystemSolve, IMPLICIT_EXTENDED_SYSTEM_BPX, synchronous (handle created with solver = 16)"""
        u = np.zeros(self.h.n[0])
        er, rr = C.c_double(0), C.c_double(0)
        self.L.ref_solve_iebpx.restype = C.c_int
        self.L.ref_solve_iebpx.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, DP, DP, DP]
        it = self.L.ref_solve_iebpx(self.handle, num_cycles, tol, mu, delta, dptr(u), C.byref(er), C.byref(rr))
        return dict(x=u, iters=it, ext_relres=er.value, relres=rr.value)

    def solve_sync_det(self, num_cycles, tol=1e-9):
        """race-free run of the reference's grouped additive cycle (see ref_driver.cpp)"""
        u = np.zeros(self.h.n[0])
        hist = np.zeros(num_cycles + 1)
        secs = C.c_double(0)
        k = self.L.ref_solve_sync_det(self.handle, num_cycles, tol, dptr(u), dptr(hist), C.byref(secs))
        return dict(u=u, hist=hist[:k + 1], cycles=k, seconds=secs.value)

    def close(self):
        if self.handle:
            self.L.ref_destroy(self.handle)
            self.handle = None
