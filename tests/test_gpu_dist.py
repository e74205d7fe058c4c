"""GPU tests of the multi-GPU path (csrc/dist.cu through the C ABI).  With one GPU: a 1-rank
communicator exercises the distributed cycle / residual / norm code against the oracle.  With >= 2 GPUs:
two processes, one per GPU, NCCL halo exchange -- the history must equal the GLOBAL oracle history."""
import os
import socket
import sys

import numpy as np
import pytest

import async_multigrid_b200 as amg
from async_multigrid_b200 import hierarchy as H, partition as PT
from conftest import HIST_TOL, assert_hist_close
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(prob, n, w, post=1):
    A = H.laplacian(prob, n)
    h = H.amg_setup(A)
    h.build_transfers(H.MULTADD, w, num_pre=1, num_post=post)
    return h, H.rand_rhs(A.nrows)


@pytest.mark.parametrize("post,direct", [(1, 0), (0, 0), (1, 1)])
def test_single_rank_communicator_matches_oracle(post, direct):
    w = 0.9
    h, b = _setup("7pt", 20, w, post)
    plan = PT.RankPlan(h, 1, 0)
    s = amg.DistSolver(plan, amg.solver.dist_unique_id(), w, num_pre=1, num_post=post, coarse_solve=bool(direct))
    s.set_rhs(b)
    hist, secs = s.solve_sync(1e-9, 100)
    _, want, _ = O.Problem(h, H.MULTADD, H.JACOBI, w, num_pre=1, num_post=post, coarse_solve=direct).solve_sync(b, 1e-9, 100)
    assert len(hist) == len(want)
    assert_hist_close(hist, want)
    u = s.get_solution()
    true = O.norm2(O.spgemv(h.A[0], u, b, -1.0, 1.0)) / O.norm2(b)
    assert abs(true - hist[-1]) <= 1e-12
    s.close()


def test_single_rank_factorised_level0_matches_oracle():
    w = 0.9
    h, b = _setup("7pt", 20, w)
    _, want, _ = O.Problem(h, H.MULTADD, H.JACOBI, w).solve_sync(b, 1e-9, 100)       # explicit products
    hf = H.Hierarchy(h.A, h.P_plain)
    hf.cpts = h.cpts
    hf.build_transfers(H.MULTADD, w, factor_level0=True)
    s = amg.DistSolver(PT.RankPlan(hf, 1, 0), amg.solver.dist_unique_id(), w, factor_level0=True)
    s.set_rhs(b)
    hist, _ = s.solve_sync(1e-9, 100)
    assert_hist_close(hist, want)
    s.close()


@pytest.mark.parametrize("accel", [1, 2])
def test_single_rank_dmem_acceleration_matches_oracle(accel):
    """DMEM_ChebyUpdate on the accumulated correction (Chebyshev / second-order Richardson)"""
    w = 0.9
    h, b = _setup("7pt", 20, w)
    pb = O.Problem(h, H.MULTADD, H.JACOBI, w)
    alpha, beta = pb.eigs_power(20)
    mu, delta = (beta + alpha) / (beta - alpha), 2.0 / (beta + alpha)
    _, want = pb.solve_sync_dmem(b, 1e-9, 100, accel, mu, delta)
    assert len(want) - 1 < 27                                   # the acceleration pays (27 cycles without)
    s = amg.DistSolver(PT.RankPlan(h, 1, 0), amg.solver.dist_unique_id(), w)
    s.set_rhs(b)
    hist, _ = s.solve_sync(1e-9, 100, accel=accel, mu=mu, delta=delta)
    assert_hist_close(hist, want)
    s.close()


@pytest.mark.parametrize("solver,w", [(H.MULTADD, 0.9), (H.BPX, 0.8)])
def test_single_rank_power_iteration_matches_oracle(solver, w):
    """DMEM_PowerMult on the partitioned path (amgb_dist_eigs_power): with a one-rank communicator and the all-ones start vector it
    is EigsPower, whose restatement is pinned by the reference's object code; a given start vector goes through as well"""
    A = H.laplacian("7pt", 16)
    h = H.amg_setup(A)
    h.build_transfers(solver, w)
    alpha, beta = O.Problem(h, solver, H.JACOBI, w).eigs_power(20)
    s = amg.DistSolver(PT.RankPlan(h, 1, 0), amg.solver.dist_unique_id(), w, solver=solver)
    mu, delta, a, bb = s.DMEM_PowerMult(20)
    assert abs(a - alpha) <= 1e-9 * abs(alpha) and abs(bb - beta) <= 1e-9 * abs(beta)
    assert abs(mu - (beta + alpha) / (beta - alpha)) <= 1e-8 * mu
    u0 = H.rand_rhs(h.n[0], 0.0, 1.0, 0) - 0.5                # src/DMEM_Eig.cpp:41: RandDouble(0.0, 1.0) - .5 after srand(rank)
    _, _, a2, b2 = s.DMEM_PowerMult(20, u0)
    assert 0.0 < a2 < b2 and abs(b2 - beta) <= 0.2 * beta      # another start vector, the same dominant eigenvalue within the 20-step accuracy
    s.close()


def test_single_rank_bpx_matches_oracle():
    """SYNC_BPX in the partitioned path (plain P, R = P^T, one Jacobi sweep on every level incl. the coarsest)"""
    w = 0.6
    A = H.laplacian("7pt", 16)
    h = H.amg_setup(A)
    h.build_transfers(H.BPX, w)
    b = H.rand_rhs(A.nrows)
    s = amg.DistSolver(PT.RankPlan(h, 1, 0), amg.solver.dist_unique_id(), w, solver=H.BPX)
    s.set_rhs(b)
    hist, _ = s.solve_sync(1e-30, 10)
    _, want, _ = O.Problem(h, H.BPX, H.JACOBI, w).solve_sync(b, 1e-30, 10)
    assert len(hist) == len(want)
    assert np.max(np.abs(hist - want) / want) <= 1e-10        # BPX alone is not convergent: relative to the growing residual
    s.close()


@pytest.mark.parametrize("solver,smoother,w,post", [(H.AFACX, H.JACOBI, 0.6, 1), (H.AFACX, H.L1_JACOBI, 0.9, 1),
                                                     (H.MULTADD, H.L1_JACOBI, 0.9, 1), (H.MULTADD, H.L1_JACOBI, 0.9, 0)])
def test_single_rank_afacx_and_l1_match_oracle(solver, smoother, w, post):
    """AFACx with the SMEM / SEQ meaning (src/SEQ_AMG.cpp:172-208) in the row-partitioned path, and the L1-Jacobi
    smoother in the partitioned path"""
    A = H.laplacian("7pt", 18)
    h = H.amg_setup(A)
    h.build_transfers(solver, w, smooth_interp_type=smoother, num_pre=1, num_post=post)
    b = H.rand_rhs(A.nrows)
    s = amg.DistSolver(PT.RankPlan(h, 1, 0), amg.solver.dist_unique_id(), w, num_pre=1, num_post=post, solver=solver, smoother=smoother)
    s.set_rhs(b)
    hist, _ = s.solve_sync(1e-9, 100)
    _, want, _ = O.Problem(h, solver, smoother, w, num_pre=1, num_post=post).solve_sync(b, 1e-9, 100)
    assert len(hist) == len(want) and hist[-1] < 1e-6          # (AFACx with L1-Jacobi needs more than 100 cycles for 1e-9)
    assert_hist_close(hist, want)
    s.close()


_LIVE = []


@pytest.fixture(autouse=True)
def _reap_workers():
    """whatever a multi-process test leaves running is killed when the test ends (pass or fail)"""
    yield
    _reap(_LIVE)
    del _LIVE[:]


def _reap(procs):
    """a worker that hangs (e.g. a deadlocked collective) must not outlive its test: it would keep spinning on the GPUs"""
    for p in procs:
        if p.is_alive():
            p.kill()
            p.join(timeout=10)


def _worker(rank, world, port, uid_q, res_q, solver=H.MULTADD, w=0.9):
    sys.path.insert(0, ROOT)
    import async_multigrid_b200 as amg2
    from async_multigrid_b200 import hierarchy as H2, partition as PT2
    A = H2.laplacian("7pt", 32)
    h = H2.amg_setup(A)
    fact = solver == H2.MULTADD
    h.build_transfers(solver, w, factor_level0=fact)      # Multadd: plain P_0 / R_0, the factorised form adds two halo exchanges
    b = H2.rand_rhs(A.nrows)
    plan = PT2.RankPlan(h, world, rank, plane=32 * 32, min_rows_per_rank=256)
    if rank == 0:
        uid = amg2.solver.dist_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    s = amg2.DistSolver(plan, uid, w, factor_level0=fact, device=rank, solver=solver)
    l0 = plan.layouts[0]
    s.set_rhs(b[l0.row_start:l0.row_start + l0.n_owned])
    hist, secs = s.solve_sync(1e-9, 100)
    u = s.get_solution()
    hb, ops = s.stats()
    res_q.put((rank, hist, u, l0.row_start, plan.num_dist, hb))
    s.close()


@pytest.mark.parametrize("solver,w", [(H.MULTADD, 0.9), (H.AFACX, 0.6)])
def test_two_gpus_match_global_oracle(solver, w):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    uid_q, res_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 0, uid_q, res_q, solver, w)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted([res_q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        _reap(procs)
    A = H.laplacian("7pt", 32)
    h = H.amg_setup(A)
    h.build_transfers(solver, w)
    b = H.rand_rhs(A.nrows)
    _, want, _ = O.Problem(h, solver, H.JACOBI, w).solve_sync(b, 1e-9, 100)
    u = np.concatenate([r[2] for r in res])
    for rank, hist, _, _, num_dist, hb in res:
        assert num_dist >= 2 and hb > 0
        assert_hist_close(hist, want)
    true = O.norm2(O.spgemv(h.A[0], u, b, -1.0, 1.0)) / O.norm2(b)
    assert abs(true - want[-1]) <= 1e-11


# ---- asynchronous solve across GPUs: a GPU plays one grid's rank group of DMEM_Add -----------------------------
def test_async_dist_single_gpu_equals_sequential_model():
    """with one GPU the level corrections are applied one after the other, each from the fresh residual: exactly the
    oracle's sequential model of the asynchronous additive iteration"""
    w = 0.9
    h, b = _setup("7pt", 16, w)
    s = amg.Solver(h, H.MULTADD, H.JACOBI, w)
    s.set_rhs(b)
    s.set_solution(None)
    s.ipc_open_peers([])
    cycles = 12
    for _ in range(cycles):
        for q in range(h.num_levels):
            s.async_dist_correct(q)
    rel = s.residual_norm() / O.norm2(b)
    u_want, counts, rel_want = O.Problem(h, H.MULTADD, H.JACOBI, w).solve_async_sequential(b, cycles)
    u = s.get_solution()
    assert np.max(np.abs(u - u_want)) <= 1e-11 * np.max(np.abs(u_want))
    assert abs(rel - rel_want) <= 1e-10
    s.close()


def _async_worker(rank, world, q_in, q_out, q_res):
    sys.path.insert(0, ROOT)
    import async_multigrid_b200 as amg2
    from async_multigrid_b200 import hierarchy as H2
    w = 0.9
    A = H2.laplacian("7pt", 24)
    h = H2.amg_setup(A)
    h.build_transfers(H2.MULTADD, w)
    b = H2.rand_rhs(A.nrows)
    s = amg2.Solver(h, H2.MULTADD, H2.JACOBI, w, device=rank)
    s.set_rhs(b)
    s.set_solution(None)
    q_out.put((rank, s.ipc_export_solution()))
    handles = q_in.get(timeout=120)                      # the other ranks' handles, from the parent
    s.ipc_open_peers(handles)
    q_out.put((rank, "ready"))
    assert q_in.get(timeout=120) == "go"
    levels = [q for q in range(h.num_levels - 1) if q % world == rank]
    for _ in range(80):
        for q in levels:
            s.async_dist_correct(q)
    s.synchronize()
    q_out.put((rank, "done"))
    assert q_in.get(timeout=120) == "all done"            # every peer's reductions have landed
    q_res.put((rank, s.residual_norm(), s.get_solution(), levels))
    assert q_in.get(timeout=120) == "close"
    s.close()


def test_async_dist_two_gpus_converge():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    world = 2
    q_in = [ctx.Queue() for _ in range(world)]
    q_out, q_res = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_async_worker, args=(r, world, q_in[r], q_out, q_res)) for r in range(world)]
    _LIVE.extend(procs)
    for p in procs:
        p.start()
    handles = dict(q_out.get(timeout=150) for _ in range(world))
    for r in range(world):
        q_in[r].put([handles[o] for o in range(world) if o != r])
    for phase, reply in (("ready", "go"), ("done", "all done")):
        got = [q_out.get(timeout=150) for _ in range(world)]
        assert all(g[1] == phase for g in got)
        for r in range(world):
            q_in[r].put(reply)
    res = sorted([q_res.get(timeout=150) for _ in range(world)], key=lambda x: x[0])
    for r in range(world):
        q_in[r].put("close")
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    h, b = _setup("7pt", 24, 0.9)
    assert sorted(res[0][3] + res[1][3]) == list(range(h.num_levels - 1))       # every working level has an owner
    assert np.array_equal(res[0][2], res[1][2]) or np.max(np.abs(res[0][2] - res[1][2])) <= 1e-12    # the same u everywhere
    true = O.norm2(O.spgemv(h.A[0], res[0][2], b, -1.0, 1.0)) / O.norm2(b)
    assert true < 1e-9, true
    assert abs(res[0][1] / O.norm2(b) - true) <= 1e-12


# ---- the asynchronous additive solve ROW-PARTITIONED over the GPUs (csrc/dist_async.cu) ---------------------------------
@pytest.mark.parametrize("fact0,smoother,w", [(True, H.JACOBI, 0.9), (False, H.JACOBI, 0.9), (True, H.L1_JACOBI, 1.0)])
def test_partitioned_async_single_rank_converges(fact0, smoother, w):
    """one rank: no exchange steps, the persistent kernel runs the single-GPU programs on the (unpartitioned) row block
    through the partitioned path's vectors -- the solve must reach the tolerance with the reported correction counts, and
    the reported relative residual must be the true one"""
    A = H.laplacian("7pt", 24)
    h = H.amg_setup(A)
    h.build_transfers(H.MULTADD, w, smooth_interp_type=smoother, factor_level0=fact0)
    b = H.rand_rhs(A.nrows)
    s = amg.DistSolver(PT.RankPlan(h, 1, 0), amg.solver.dist_unique_id(), w, factor_level0=fact0, smoother=smoother)
    s.set_rhs(b)
    K = 60
    cor, rel, secs = s.DMEM_Add_async(K)
    assert list(cor) == [K] * h.num_levels
    assert rel < 1e-8, rel
    u = s.get_solution()
    true = O.norm2(O.spgemv(h.A[0], u, b, -1.0, 1.0)) / O.norm2(b)
    assert abs(true - rel) <= 1e-12
    cb, t = s.async_groups()
    assert cb[0] == 0 and all(cb[q + 1] > cb[q] for q in range(h.num_levels)) and secs > 0
    s.close()


def _part_async_worker(rank, world, uid_q, res_q, n, K, solver, w):
    sys.path.insert(0, ROOT)
    import async_multigrid_b200 as amg2
    from async_multigrid_b200 import hierarchy as H2, partition as PT2
    A = H2.laplacian("7pt", n)
    h = H2.amg_setup(A)
    fact = solver == H2.MULTADD
    h.build_transfers(solver, w, factor_level0=fact)
    b = H2.rand_rhs(A.nrows)
    plan = PT2.RankPlan(h, world, rank, plane=n * n, min_rows_per_rank=256)
    if rank == 0:
        uid = amg2.solver.dist_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    s = amg2.DistSolver(plan, uid, w, factor_level0=fact, device=rank, solver=solver)
    l0 = plan.layouts[0]
    s.set_rhs(b[l0.row_start:l0.row_start + l0.n_owned])
    cor, rel, secs = s.DMEM_Add_async(K)
    u = s.get_solution()
    hb, _ = s.stats()
    # a second solve from the solution of the first (the flags and group copies are re-armed per launch)
    cor2, rel2, _ = s.DMEM_Add_async(5)
    res_q.put((rank, list(cor), rel, u, plan.num_dist, hb, secs, rel2))
    s.close()


@pytest.mark.parametrize("solver,w,K,bound", [(H.MULTADD, 0.9, 60, 1e-8), (H.AFACX, 0.6, 90, 1e-4)])
def test_partitioned_async_two_gpus_converge(solver, w, K, bound):
    """two GPUs, one z-slab each: every level group exchanges its boundaries with the same group on the other GPU (stores
    over NVLink + step flags), the groups run asynchronously; to 1e-8 with every group's K corrections, the same global
    residual on both ranks, and it is the true one"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    n = 32
    uid_q, res_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_part_async_worker, args=(r, 2, uid_q, res_q, n, K, solver, w)) for r in range(2)]
    _LIVE.extend(procs)
    for p in procs:
        p.start()
    res = sorted([res_q.get(timeout=150) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A = H.laplacian("7pt", n)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    u = np.concatenate([r[3] for r in res])
    true = O.norm2(O.spgemv(A, u, b, -1.0, 1.0)) / O.norm2(b)
    for rank, cor, rel, _, num_dist, hb, secs, rel2 in res:
        assert cor == [K] * h.num_levels
        assert num_dist >= 2 and hb > 0 and secs > 0
        assert abs(rel - true) <= 1e-12
        assert np.isfinite(rel2) and rel2 < 10.0            # relative to the (already tiny) residual the second solve started from
    assert true < bound, true


# ---- DMEM_AsyncSmooth: asynchronous (L1-)Jacobi on the fine grid across GPUs (src/DMEM_Smooth.cpp:16-313) -------------
@pytest.mark.parametrize("smoother,kind", [(H.JACOBI, "jacobi"), (H.L1_JACOBI, "l1_jacobi")])
def test_async_smooth_single_rank_equals_jacobi_sweeps(smoother, kind):
    """one rank has no neighbour to be late: `sweeps` relaxations are exactly `sweeps` (L1-)Jacobi sweeps of the oracle"""
    w = 0.9
    A = H.laplacian("7pt", 14)
    h = H.amg_setup(A)
    h.build_transfers(H.MULTADD, w, smooth_interp_type=smoother)
    b = H.rand_rhs(A.nrows)
    s = amg.DistSolver(PT.RankPlan(h, 1, 0), amg.solver.dist_unique_id(), w, smoother=smoother)
    s.set_rhs(b)
    s.zero_solution()
    s.ipc_open_neighbours(None, None)
    s.DMEM_AsyncSmooth(17)
    s.synchronize()
    u = s.get_solution()
    want = O.smooth(kind, h.A[0], b, w, sweeps=17, zero_flag=1, l1=h.l1_norms()[0])
    assert np.max(np.abs(u - want)) <= 1e-13 * np.max(np.abs(want))
    rn = s.residual_norm()
    assert abs(rn - O.norm2(O.spgemv(h.A[0], u, b, -1.0, 1.0))) <= 1e-12 * O.norm2(b)
    s.close()


def _async_smooth_worker(rank, world, q_in, q_out, q_res):
    sys.path.insert(0, ROOT)
    import async_multigrid_b200 as amg2
    from async_multigrid_b200 import hierarchy as H2, partition as PT2
    w = 0.9
    A = H2.laplacian("7pt", 24)
    h = H2.amg_setup(A)
    h.build_transfers(H2.MULTADD, w)
    b = H2.rand_rhs(A.nrows)
    plan = PT2.RankPlan(h, world, rank, plane=24 * 24, min_rows_per_rank=256)
    if rank == 0:
        uid = amg2.solver.dist_unique_id()
        q_out.put((rank, "uid", uid))
    uid = q_in.get(timeout=120)
    s = amg2.DistSolver(plan, uid, w, device=rank)
    l0 = plan.layouts[0]
    s.set_rhs(b[l0.row_start:l0.row_start + l0.n_owned])
    s.zero_solution()
    q_out.put((rank, "handle", s.ipc_export_solution()))
    handles = q_in.get(timeout=120)
    s.ipc_open_neighbours(handles.get(rank - 1), handles.get(rank + 1))
    r0 = s.residual_norm()                                # collective: both ranks are set up past this point
    sweeps = 300 + 100 * rank                             # LOCAL stop rule: the ranks do different amounts of work
    s.DMEM_AsyncSmooth(sweeps)
    s.synchronize()
    q_out.put((rank, "done", None))
    assert q_in.get(timeout=120) == "all done"            # every neighbour's stores have landed
    rn = s.residual_norm()
    q_res.put((rank, r0, rn, s.get_solution(), s.stats()[0]))
    assert q_in.get(timeout=120) == "close"
    s.close()


def test_async_smooth_two_gpus():
    """two GPUs relax their slabs without waiting for each other (300 and 400 sweeps).  Any interleaving is a legal
    outcome: lock-step gives the residual of 300-400 synchronous Jacobi sweeps (0.14 - 0.07 here), one rank running
    entirely before the other 4.0 (of ||b|| = 67.8); the result must lie in that range, and the norm the solver reports
    must be the true global one"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    world = 2
    q_in = [ctx.Queue() for _ in range(world)]
    q_out, q_res = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_async_smooth_worker, args=(r, world, q_in[r], q_out, q_res)) for r in range(world)]
    _LIVE.extend(procs)
    for p in procs:
        p.start()
    _, tag, uid = q_out.get(timeout=150)
    assert tag == "uid"
    for r in range(world):
        q_in[r].put(uid)
    handles = {}
    for _ in range(world):
        r, tag, hd = q_out.get(timeout=150)
        assert tag == "handle"
        handles[r] = hd
    for r in range(world):
        q_in[r].put(handles)
    got = [q_out.get(timeout=150) for _ in range(world)]
    assert all(g[1] == "done" for g in got)
    for r in range(world):
        q_in[r].put("all done")
    res = sorted([q_res.get(timeout=150) for _ in range(world)], key=lambda x: x[0])
    for r in range(world):
        q_in[r].put("close")
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A = H.laplacian("7pt", 24)
    b = H.rand_rhs(A.nrows)
    u = np.concatenate([r[3] for r in res])
    true = O.norm2(O.spgemv(A, u, b, -1.0, 1.0))
    assert abs(res[0][2] - true) <= 1e-10 * O.norm2(b) and abs(res[1][2] - true) <= 1e-10 * O.norm2(b)
    assert abs(res[0][1] - O.norm2(b)) <= 1e-12 * O.norm2(b)
    r400 = O.norm2(O.spgemv(A, O.smooth("jacobi", A, b, 0.9, sweeps=400), b, -1.0, 1.0))
    assert 0.5 * r400 <= true <= 0.1 * O.norm2(b), (r400, true, O.norm2(b))
    assert res[0][4] > 0 and res[1][4] > 0                 # boundary values really went over NVLink


def test_single_rank_graph_captured_cycle_matches_oracle():
    """the partitioned cycle replayed as a CUDA graph (NCCL calls captured) must give the history of the eager path"""
    w = 0.9
    h, b = _setup("7pt", 20, w)
    _, want, _ = O.Problem(h, H.MULTADD, H.JACOBI, w).solve_sync(b, 1e-9, 100)
    os.environ["AMGB_DIST_GRAPH"] = "1"
    try:
        s = amg.DistSolver(PT.RankPlan(h, 1, 0), amg.solver.dist_unique_id(), w)
    finally:
        del os.environ["AMGB_DIST_GRAPH"]
    s.set_rhs(b)
    for _ in range(2):                      # second solve replays the instantiated graph
        hist, _ = s.solve_sync(1e-9, 100)
        assert_hist_close(hist, want)
    s.close()
