"""Planning of the row-partitioned asynchronous solve (csrc/dist_async.cu dist_async_plan, through the host-only probe
amgb_dist_async_plan), interpreted on the CPU for 2 and 3 ranks (tests/dist_async_emulator.py):

  * in the lock-step interleaving every ghost value is current, so the partitioned plans must reproduce the single-GPU
    programs interpreted on the unpartitioned hierarchy (tests/async_emulator.py, itself checked against the oracle's
    sequential model of SMEM_Async_Add_AMG) to rounding -- slots, element offsets, boundary pushes and the gather into the
    replicated tail are all exercised by that;
  * under random interleavings (groups drifting apart, as on the device; the exchange flags keep the ranks of one group in
    step) the iteration converges like the single-GPU programs under random interleavings, and never deadlocks;
  * every push lands inside a ghost range (or a peer's copy of a replicated vector), never on owned entries."""
import numpy as np
import pytest

import async_multigrid_b200 as amg  # noqa: F401
from async_multigrid_b200 import hierarchy as H
from async_multigrid_b200 import solver as S
from async_emulator import Emulator
from dist_async_emulator import DistAsyncEmulator, PUSH, X, Y


def _problem(prob="7pt", n=12, solver=H.MULTADD, w=0.9, post=1, fact0=False, smoother=H.JACOBI):
    A = H.laplacian(prob, n)
    h = H.amg_setup(A)
    assert h.num_levels >= 3
    if fact0:
        h.build_transfers(solver, w, smooth_interp_type=smoother, num_pre=1, num_post=post)
        hf = H.Hierarchy(h.A, h.P_plain)
        hf.cpts = h.cpts
        hf.build_transfers(solver, w, smooth_interp_type=smoother, factor_level0=True)
        h = hf
    else:
        h.build_transfers(solver, w, smooth_interp_type=smoother, num_pre=1, num_post=post)
    return h, H.rand_rhs(A.nrows), (n * n if prob != "5pt" else n)


CASES = [
    # prob, n, ranks, min_rows, solver, async solver, w, post, fact0, smoother
    ("7pt", 12, 2, 40, H.MULTADD, H.ASYNC_MULTADD, 0.9, 1, False, H.JACOBI),
    ("7pt", 12, 3, 40, H.MULTADD, H.ASYNC_MULTADD, 0.9, 1, True, H.JACOBI),
    ("7pt", 12, 2, 40, H.MULTADD, H.ASYNC_MULTADD, 0.9, 1, True, H.JACOBI),
    ("5pt", 32, 3, 30, H.MULTADD, H.ASYNC_MULTADD, 0.8, 0, False, H.JACOBI),
    ("5pt", 32, 2, 30, H.MULTADD, H.ASYNC_MULTADD, 1.0, 1, False, H.L1_JACOBI),
    ("7pt", 12, 2, 40, H.AFACX, H.ASYNC_AFACX, 0.6, 1, False, H.JACOBI),
    ("7pt", 12, 2, 10 ** 9, H.MULTADD, H.ASYNC_MULTADD, 0.9, 1, True, H.JACOBI),      # level 0 alone is partitioned
    ("7pt", 32, 4, 40, H.MULTADD, H.ASYNC_MULTADD, 0.9, 1, True, H.JACOBI),           # ranks with two neighbours, two partitioned levels
    ("7pt", 16, 8, 40, H.MULTADD, H.ASYNC_MULTADD, 0.9, 1, True, H.JACOBI),           # eight ranks: the gather has seven destinations
]


@pytest.mark.parametrize("prob,n,ranks,min_rows,solver,asolver,w,post,fact0,smoother", CASES)
def test_lockstep_equals_the_single_gpu_programs(prob, n, ranks, min_rows, solver, asolver, w, post, fact0, smoother):
    h, b, plane = _problem(prob, n, solver, w, post, fact0, smoother)
    sym = solver == H.MULTADD and post > 0
    K = 5
    ref = Emulator(h, S.async_program(h.num_levels, asolver, smoother, symmetric=sym, factor_level0=fact0), b, smoother, w)
    want = ref.run(K)
    em = DistAsyncEmulator(h, ranks, b, asolver, smoother, w, symmetric=sym, factor_level0=fact0, plane=plane, min_rows_per_rank=min_rows)
    assert em.num_dist >= 1
    got = em.run_lockstep(K)
    assert em.pushed > 0
    assert all(em.count[p] == [K] * h.num_levels for p in range(ranks))
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    assert abs(em.relres() - ref.relres()) <= 1e-12


def test_partition_has_distributed_and_replicated_levels():
    """the cases above must cover both kinds of level and the gather between them"""
    h, b, plane = _problem("7pt", 12)
    em = DistAsyncEmulator(h, 2, b, H.ASYNC_MULTADD, H.JACOBI, 0.9, plane=plane, min_rows_per_rank=40)
    assert 2 <= em.num_dist < h.num_levels
    lay = [pl.layouts for pl in em.plans]
    gathers = 0
    for p in range(2):
        for q in range(h.num_levels):
            for op in em.progs[p][q]:
                if op.type != PUSH or op.count == 0:
                    continue
                s = op.slot[Y]
                assert s == op.slot[X] and em.slot_group[s] == q
                vid = int(em.slot_vec[s])
                lvl = 0 if vid // 64 in (7, 8, 9) else vid % 64
                dst = lay[op.dst_rank][lvl]
                src = lay[p][lvl]
                if dst.distributed:
                    # a boundary push lands exactly on the neighbour's ghost range facing this rank
                    if op.dst_rank == p - 1:
                        assert op.elem[Y] == dst.halo_lo + dst.n_owned and op.count == dst.halo_hi
                        assert op.elem[X] == src.halo_lo
                    else:
                        assert op.dst_rank == p + 1 and op.elem[Y] == 0 and op.count == dst.halo_lo
                        assert op.elem[X] == src.halo_lo + src.n_owned - op.count
                else:
                    gathers += 1
                    assert lvl == em.num_dist and op.elem[X] == op.elem[Y] == src.row_start and op.count == src.n_owned
    assert gathers > 0


@pytest.mark.parametrize("ranks,fact0,seed", [(2, True, 1), (3, True, 2), (3, False, 3), (4, True, 4)])
def test_random_interleaving_converges_like_one_gpu(ranks, fact0, seed):
    h, b, plane = _problem("7pt", 12, fact0=fact0)
    em = DistAsyncEmulator(h, ranks, b, H.ASYNC_MULTADD, H.JACOBI, 0.9, factor_level0=fact0, plane=plane, min_rows_per_rank=40)
    em.run_random(40, seed=seed)
    assert all(em.count[p] == [40] * h.num_levels for p in range(ranks))
    # the same programs on ONE rank under the same kind of random group interleaving: the inherent cost of asynchrony
    one = []
    for s in range(3):
        e1 = DistAsyncEmulator(h, 1, b, H.ASYNC_MULTADD, H.JACOBI, 0.9, factor_level0=fact0, plane=plane)
        e1.run_random(40, seed=100 + s)
        one.append(e1.relres())
    assert em.relres() < 1e-5, em.relres()
    assert em.relres() < 30.0 * max(one), (em.relres(), one)


def test_eight_ranks_never_deadlock_under_random_interleaving():
    """eight ranks, one partitioned level: every iteration of a deep group has one all-to-all exchange step (the gather into
    the replicated tail) between neighbour-only steps; whatever the interleaving, some group can always advance"""
    h, b, plane = _problem("7pt", 16, fact0=True)
    for seed in (11, 12):
        em = DistAsyncEmulator(h, 8, b, H.ASYNC_MULTADD, H.JACOBI, 0.9, factor_level0=True, plane=plane, min_rows_per_rank=40)
        em.run_random(30, seed=seed, max_burst=9)
        assert all(em.count[p] == [30] * h.num_levels for p in range(8))
        assert em.relres() < 1e-3, em.relres()


def test_without_the_exchange_flags_stale_ghosts_cost_orders_of_magnitude():
    """why the flags exist: the same plans with the waits ignored (a group reads whatever its ghost slots hold: restricted
    residuals and corrections of the neighbour's previous iteration)"""
    h, b, plane = _problem("7pt", 12, fact0=True)
    mk = lambda: DistAsyncEmulator(h, 3, b, H.ASYNC_MULTADD, H.JACOBI, 0.9, factor_level0=True, plane=plane, min_rows_per_rank=40)
    with_flags, without = [], []
    for seed in (1, 2):
        e = mk()
        e.run_random(40, seed=seed)
        with_flags.append(e.relres())
        e = mk()
        e.run_random(40, seed=seed, honour_waits=False)
        without.append(e.relres())
    assert min(without) > 10.0 * max(with_flags), (with_flags, without)


@pytest.mark.parametrize("ranks,fact0", [(2, True), (3, False)])
def test_direct_coarse_solve_lockstep_equals_the_single_gpu_programs(ranks, fact0):
    """DMEM's convention for the coarsest level (direct solve) in the partitioned plans: the coarsest level is replicated, its
    group restricts through the partitioned levels, gathers, solves, and prolongs back"""
    h, b, plane = _problem("7pt", 12, fact0=fact0)
    K = 5
    ref = Emulator(h, S.async_program(h.num_levels, H.ASYNC_MULTADD, H.JACOBI, symmetric=True, factor_level0=fact0, coarse_solve=True),
                   b, H.JACOBI, 0.9)
    want = ref.run(K)
    em = DistAsyncEmulator(h, ranks, b, H.ASYNC_MULTADD, H.JACOBI, 0.9, factor_level0=fact0, plane=plane, min_rows_per_rank=40,
                           coarse_solve=True)
    got = em.run_lockstep(K)
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    em2 = DistAsyncEmulator(h, ranks, b, H.ASYNC_MULTADD, H.JACOBI, 0.9, factor_level0=fact0, plane=plane, min_rows_per_rank=40,
                            coarse_solve=True)
    em2.run_random(30, seed=3)
    assert em2.relres() < 1e-4


def test_unsupported_options_are_refused():
    h, b, plane = _problem("7pt", 12)
    with pytest.raises(S.AmgError):
        DistAsyncEmulator(h, 2, b, H.BPX, H.JACOBI, 0.9, plane=plane, min_rows_per_rank=40)
    with pytest.raises(S.AmgError):
        DistAsyncEmulator(h, 2, b, H.ASYNC_MULTADD, H.HYBRID_JACOBI_GAUSS_SEIDEL, 0.9, plane=plane, min_rows_per_rank=40)


def test_single_rank_plan_has_no_pushes():
    h, b, plane = _problem("7pt", 12)
    em = DistAsyncEmulator(h, 1, b, H.ASYNC_MULTADD, H.JACOBI, 0.9, plane=plane)
    assert em.num_dist == 0
    assert not any(op.type == PUSH for q in range(h.num_levels) for op in em.progs[0][q])
    ref = Emulator(h, S.async_program(h.num_levels, H.ASYNC_MULTADD, H.JACOBI, symmetric=True), b, H.JACOBI, 0.9)
    assert np.max(np.abs(em.run_lockstep(4) - ref.run(4))) <= 1e-13
