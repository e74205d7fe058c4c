"""SELL-U ("uniform slices", amgb_options.sell_uniform): the host-side encoder of csrc/context.cu checked on the CPU against
plain CSR products -- the encoding must be lossless (same operator, a row's terms in another order) for the stencil
matrices it compresses and must leave matrices with all-different values alone.  The device kernel that consumes it
(sell_rows_team in csrc/kernels.cuh) is emulated here group by group; the GPU suite checks the kernel itself."""
import ctypes as C

import numpy as np
import pytest

import async_multigrid_b200 as amg
from async_multigrid_b200 import hierarchy as H


def encode(m, sv):
    L = amg.solver.load_library()
    IP, UP, DP = C.POINTER(C.c_int), C.POINTER(C.c_uint), C.POINTER(C.c_double)
    L.amgb_sellu_encode_host.argtypes = [C.c_int, IP, IP, DP, DP, C.POINTER(IP), C.POINTER(IP), C.POINTER(UP), C.POINTER(DP),
                                         C.POINTER(DP), IP]
    L.amgb_host_free.argtypes = [C.c_void_p]
    L.amgb_host_free.restype = None
    desc, dl, mk, gv, gs, ng = IP(), IP(), UP(), DP(), DP(), C.c_int(0)
    sv = np.ascontiguousarray(sv, dtype=np.float64)
    slices = L.amgb_sellu_encode_host(m.nrows, m.indptr.ctypes.data_as(IP), m.indices.ctypes.data_as(IP), m.data.ctypes.data_as(DP),
                                      sv.ctypes.data_as(DP), C.byref(desc), C.byref(dl), C.byref(mk), C.byref(gv), C.byref(gs), C.byref(ng))
    n = max(ng.value, 1)
    out = (np.ctypeslib.as_array(desc, shape=(max(slices, 1), 2))[:slices].copy(), np.ctypeslib.as_array(dl, shape=(n,))[:ng.value].copy(),
           np.ctypeslib.as_array(mk, shape=(n,))[:ng.value].copy(), np.ctypeslib.as_array(gv, shape=(n,))[:ng.value].copy(),
           np.ctypeslib.as_array(gs, shape=(n,))[:ng.value].copy())
    for p in (desc, dl, mk, gv, gs):
        L.amgb_host_free(p)
    return out


def emulate(m, vals, desc, dl, mk, gvals, x):
    """what sell_rows_team computes: encoded slices group by group (slice s owns groups desc[s,0] .. desc[s,0] + desc[s,1] of the
    deduplicated table), the others entry by entry"""
    y = np.zeros(m.nrows)
    for s in range(len(desc)):
        rows = range(32 * s, min(32 * s + 32, m.nrows))
        if desc[s, 1] > 0:
            for g in range(desc[s, 0], desc[s, 0] + desc[s, 1]):
                for r in rows:
                    if (int(mk[g]) >> (r & 31)) & 1:
                        y[r] += gvals[g] * x[r + dl[g]]
        else:
            for r in rows:
                for p in range(m.indptr[r], m.indptr[r + 1]):
                    y[r] += vals[p] * x[m.indices[p]]
    return y


@pytest.mark.parametrize("prob,n", [("7pt", 11), ("27pt", 7), ("5pt", 37)])
def test_stencils_are_encoded_losslessly(prob, n):
    A = H.laplacian(prob, n)
    ws = 0.9 / A.diagonal()
    sv = A.data * ws[A.indices]
    goff, dl, mk, gv, gs = encode(A, sv)
    slices = (A.nrows + 31) // 32
    per = goff[:, 1]
    assert (per > 0).sum() >= 0.9 * slices                       # nearly every slice qualifies
    width = {"7pt": 7, "27pt": 27, "5pt": 5}[prob]
    assert per.max() <= 2 * width                                  # a handful of groups per slice
    assert len(dl) <= per.sum()                                    # repeated lists are kept once (dedup checked in test_dedup below)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(A.nrows)
    S = A.to_scipy()
    np.testing.assert_allclose(emulate(A, A.data, goff, dl, mk, gv, x), S @ x, rtol=0, atol=1e-13)
    np.testing.assert_allclose(emulate(A, sv, goff, dl, mk, gs, x), S @ (ws * x), rtol=0, atol=1e-13)
    # every real entry sits in exactly one group of its slice's list
    bits = np.array([bin(int(v)).count("1") for v in mk])
    covered = sum(int(bits[goff[s, 0]:goff[s, 0] + goff[s, 1]].sum()) for s in range(slices))
    assert covered == A.nnz - sum(int(A.indptr[min(32 * s + 32, A.nrows)] - A.indptr[32 * s]) for s in range(slices) if per[s] == 0)


def test_l1_scaling_and_nonsymmetric_stencil():
    A = H.difconv(9, 8, 7, a=(40.0, -20.0, 10.0), atype=3)
    l1 = np.bincount(np.repeat(np.arange(A.nrows), np.diff(A.indptr)), weights=np.abs(A.data), minlength=A.nrows)
    sv = A.data / l1[A.indices]
    goff, dl, mk, gv, gs = encode(A, sv)
    x = np.random.default_rng(2).standard_normal(A.nrows)
    S = A.to_scipy()
    np.testing.assert_allclose(emulate(A, A.data, goff, dl, mk, gv, x), S @ x, rtol=0, atol=1e-12)
    np.testing.assert_allclose(emulate(A, sv, goff, dl, mk, gs, x), S @ (x / l1), rtol=0, atol=1e-12)


def test_galerkin_operator_is_left_alone():
    A = H.laplacian("7pt", 12)
    h = H.amg_setup(A)
    A1 = h.A[1]
    goff, dl, mk, gv, gs = encode(A1, A1.data)
    assert (goff[:, 1] > 0).sum() <= 0.2 * ((A1.nrows + 31) // 32)     # values all differ: nothing to share
    x = np.random.default_rng(3).standard_normal(A1.nrows)
    np.testing.assert_allclose(emulate(A1, A1.data, goff, dl, mk, gv, x), A1.to_scipy() @ x, rtol=0, atol=1e-12)


def test_duplicate_column_in_a_row_is_not_encoded():
    ip = np.array([0, 3, 5], dtype=np.int32)
    ix = np.array([0, 1, 1, 1, 0], dtype=np.int32)              # row 0 holds column 1 twice
    dv = np.array([2.0, -1.0, -1.0, 2.0, -1.0])
    m = H.CSR(2, 2, ip, ix, dv)
    goff, dl, mk, gv, gs = encode(m, dv)
    assert goff[:, 1].sum() == 0


def test_dedup_of_repeated_group_lists():
    """a constant-coefficient stencil on a grid whose lines are whole slices has one group list per boundary pattern"""
    A = H.laplacian("7pt", 32)                                    # 32 rows per grid line = one slice per line
    goff, dl, mk, gv, gs = encode(A, A.data)
    assert goff[:, 1].min() > 0
    assert len(dl) <= 9 * 7                                       # interior / low / high in y and z: 9 patterns of <= 7 groups
    x = np.random.default_rng(5).standard_normal(A.nrows)
    np.testing.assert_allclose(emulate(A, A.data, goff, dl, mk, gv, x), A.to_scipy() @ x, rtol=0, atol=1e-13)
