"""The per-group programs of the persistent asynchronous kernel (csrc/async.cu async_build_program), interpreted on the CPU
(tests/async_emulator.py) against the oracle: what the host hands the device kernel must BE the reference's asynchronous
iteration.  With one working group (two-level hierarchy) the chaotic iteration is deterministic and equals the oracle's
sequential model (itself pinned bit for bit by SMEM_Async_Add_AMG's object code, tests/golden/async_two_level.npz); with
several groups taking turns every variant must converge, and variants that are algebraically the same operator (explicit /
factorised level-0 transfers) must agree to rounding under the same interleaving."""
import numpy as np
import pytest

import async_multigrid_b200 as amg
from async_multigrid_b200 import hierarchy as H
from async_multigrid_b200 import solver as S
from oracle import oracle as O
from async_emulator import Emulator


def _two_level(prob="5pt", n=20, solver=H.MULTADD, w=0.9, post=1):
    A = H.laplacian(prob, n)
    h = H.amg_setup(A, max_levels=2)
    h.build_transfers(solver, w, num_pre=1, num_post=post)
    return h, H.rand_rhs(A.nrows)


@pytest.mark.parametrize("solver,base,w,post,sweeps", [(H.ASYNC_MULTADD, H.MULTADD, 0.9, 1, 1), (H.ASYNC_MULTADD, H.MULTADD, 0.9, 1, 2),
                                                       (H.ASYNC_MULTADD, H.MULTADD, 0.8, 0, 1), (H.ASYNC_MULTADD, H.MULTADD, 0.8, 0, 3),
                                                       (H.ASYNC_AFACX, H.AFACX, 0.6, 1, 1), (H.ASYNC_AFACX, H.AFACX, 0.6, 1, 2)])
def test_single_group_program_is_the_sequential_model(solver, base, w, post, sweeps):
    h, b = _two_level(solver=base, w=w, post=post)
    sym = base == H.MULTADD and post > 0
    prog = S.async_program(2, solver, H.JACOBI, symmetric=sym, fine_sweeps=sweeps, coarse_sweeps=sweeps)
    for K in (1, 6):
        e = Emulator(h, prog, b, H.JACOBI, w)
        u = e.run(K)
        want, counts, rel = O.Problem(h, base, H.JACOBI, w, num_pre=1, num_post=post, fine_sweeps=sweeps,
                                      coarse_sweeps=sweeps).solve_async_sequential(b, K)
        assert e.count == [K, K]
        assert np.max(np.abs(u - want)) <= 1e-13 * np.max(np.abs(want))
        assert abs(e.relres() - rel) <= 1e-13


def test_single_group_read_res_program_is_the_sequential_model():
    h, b = _two_level()
    prog = S.async_program(2, H.ASYNC_MULTADD, H.JACOBI, symmetric=True, read_type=1)
    e = Emulator(h, prog, b, H.JACOBI, 0.9)
    u = e.run(9, read_res=True)
    want, counts, rel = O.Problem(h, H.MULTADD, H.JACOBI, 0.9).solve_async_sequential(b, 9, read_res=True)
    assert np.max(np.abs(u - want)) <= 1e-12 * np.max(np.abs(want))


def test_l1_jacobi_single_group():
    h, b = _two_level("7pt", 9)
    prog = S.async_program(2, H.ASYNC_MULTADD, H.L1_JACOBI, symmetric=True)
    e = Emulator(h, prog, b, H.L1_JACOBI, 1.0)
    # transfers smoothed with the L1 factors
    h.build_transfers(H.MULTADD, 1.0, smooth_interp_type=H.L1_JACOBI)
    e = Emulator(h, prog, b, H.L1_JACOBI, 1.0)
    u = e.run(5)
    want, _, _ = O.Problem(h, H.MULTADD, H.L1_JACOBI, 1.0).solve_async_sequential(b, 5)
    assert np.max(np.abs(u - want)) <= 1e-13 * np.max(np.abs(want))


def _multi(n=10):
    A = H.laplacian("7pt", n)
    h = H.amg_setup(A)
    assert h.num_levels >= 3
    return A, h, H.rand_rhs(A.nrows)


def test_factorised_level0_program_equals_explicit_products():
    A, h, b = _multi()
    w = 0.9
    h.build_transfers(H.MULTADD, w)
    pe = S.async_program(h.num_levels, H.ASYNC_MULTADD, H.JACOBI, symmetric=True)
    ue = Emulator(h, pe, b, H.JACOBI, w).run(7)
    hf = H.Hierarchy(h.A, h.P_plain)
    hf.build_transfers(H.MULTADD, w, factor_level0=True)
    pf = S.async_program(h.num_levels, H.ASYNC_MULTADD, H.JACOBI, symmetric=True, factor_level0=True)
    uf = Emulator(hf, pf, b, H.JACOBI, w).run(7)
    assert np.max(np.abs(ue - uf)) <= 1e-12 * np.max(np.abs(ue))


@pytest.mark.parametrize("kw", [dict(), dict(read_type=1), dict(async_type=1), dict(res_compute_type=1),
                                dict(res_compute_type=1, factor_level0=True), dict(factor_level0=True, read_type=1)])
def test_every_variant_converges_when_the_groups_take_turns(kw):
    A, h, b = _multi()
    w = 0.9
    fact0 = kw.get("factor_level0", False)
    h.build_transfers(H.MULTADD, w, factor_level0=fact0)
    prog = S.async_program(h.num_levels, H.ASYNC_MULTADD, H.JACOBI, symmetric=True, **kw)
    e = Emulator(h, prog, b, H.JACOBI, w, first_group=1 if kw.get("res_compute_type") else 0)
    K = 100 if kw.get("res_compute_type") else 60     # (global: every row of level 0 is smoothed once per round only)
    e.run(K, read_res=bool(kw.get("read_type")))
    assert e.relres() < 1e-9, e.relres()
    first = 1 if kw.get("res_compute_type") else 0
    assert e.count[first:] == [K] * (h.num_levels - first)


def test_unsupported_combinations_are_refused():
    with pytest.raises(S.AmgError):
        S.async_program(3, H.ASYNC_AFACX, H.JACOBI, symmetric=False, res_compute_type=1)       # src/SMEM_Main.cpp:650-660
    with pytest.raises(S.AmgError):
        S.async_program(3, H.ASYNC_MULTADD, H.JACOBI, symmetric=True, async_type=1, read_type=1)
    with pytest.raises(S.AmgError):
        S.async_program(3, H.ASYNC_AFACX, H.JACOBI, symmetric=False, factor_level0=True)


@pytest.mark.parametrize("fact0", [False, True])
def test_direct_coarse_solve_program_is_the_dmem_sequential_model(fact0):
    """amgb_options.coarse_solve: the coarsest group solves its level directly (DMEM's AddCycle, src/DMEM_Add.cpp:262-264).  The
    oracle's sequential model of the DMEM convention (pinned by AddCycle's object code run for every grid in turn,
    tests/golden/dmem.npz) gives every grid the FRESH fine residual; handing every group that residual before its turn, the
    programs' chains -- restriction down to the coarsest level, the direct solve, prolongation, update -- must reproduce it"""
    from async_emulator import V_R, V_UL
    A, h, b = _multi()
    w = 0.9
    h.build_transfers(H.MULTADD, w)
    K = 6
    want, counts, rel = O.Problem(h, H.MULTADD, H.JACOBI, w, coarse_solve=1).solve_async_sequential(b, K)
    hh = h
    if fact0:
        hh = H.Hierarchy(h.A, h.P_plain)
        hh.build_transfers(H.MULTADD, w, factor_level0=True)
    prog = S.async_program(h.num_levels, H.ASYNC_MULTADD, H.JACOBI, symmetric=True, factor_level0=fact0, coarse_solve=True)
    assert len(prog[-1]) > 1                                   # the coarsest group works
    e = Emulator(hh, prog, b, H.JACOBI, w)
    for _ in range(K):
        for q in range(h.num_levels):
            e.vec(q, V_R * 64)[:] = e.f - e.A[0] @ e.u
            e.vec(q, V_UL * 64)[:] = e.u
            e.run_group_iteration(q)
    assert e.count == [K] * h.num_levels
    assert np.max(np.abs(e.u - want)) <= 1e-12 * np.max(np.abs(want))
    assert abs(e.relres() - rel) <= 1e-12
    # ... and with the groups simply taking turns (every group on its own, older residual) the solve converges as with the
    # SMEM convention (on this grid the coarsest level has a handful of points and contributes little either way)
    e1 = Emulator(hh, prog, b, H.JACOBI, w)
    e1.run(12)
    e0 = Emulator(hh, S.async_program(h.num_levels, H.ASYNC_MULTADD, H.JACOBI, symmetric=True, factor_level0=fact0), b, H.JACOBI, w)
    e0.run(12)
    assert e1.relres() < 1e-3 and e1.relres() < 1.5 * e0.relres()


def test_direct_coarse_solve_is_refused_with_the_global_residual():
    with pytest.raises(S.AmgError):
        S.async_program(4, H.ASYNC_MULTADD, H.JACOBI, symmetric=True, res_compute_type=1, coarse_solve=True)
