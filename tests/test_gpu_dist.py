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
from conftest import HIST_TOL
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
    assert np.max(np.abs(hist - want)) <= HIST_TOL
    u = s.get_solution()
    true = O.norm2(O.spgemv(h.A[0], u, b, -1.0, 1.0)) / O.norm2(b)
    assert abs(true - hist[-1]) <= 1e-12
    s.close()


def _worker(rank, world, port, uid_q, res_q):
    sys.path.insert(0, ROOT)
    import async_multigrid_b200 as amg2
    from async_multigrid_b200 import hierarchy as H2, partition as PT2
    w = 0.9
    A = H2.laplacian("7pt", 32)
    h = H2.amg_setup(A)
    h.build_transfers(H2.MULTADD, w)
    b = H2.rand_rhs(A.nrows)
    plan = PT2.RankPlan(h, world, rank, plane=32 * 32, min_rows_per_rank=256)
    if rank == 0:
        uid = amg2.solver.dist_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    s = amg2.DistSolver(plan, uid, w, device=rank)
    l0 = plan.layouts[0]
    s.set_rhs(b[l0.row_start:l0.row_start + l0.n_owned])
    hist, secs = s.solve_sync(1e-9, 100)
    u = s.get_solution()
    hb, ops = s.stats()
    res_q.put((rank, hist, u, l0.row_start, plan.num_dist, hb))
    s.close()


def test_two_gpus_match_global_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    uid_q, res_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 0, uid_q, res_q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([res_q.get(timeout=300) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    h, b = _setup("7pt", 32, 0.9)
    _, want, _ = O.Problem(h, H.MULTADD, H.JACOBI, 0.9).solve_sync(b, 1e-9, 100)
    u = np.concatenate([r[2] for r in res])
    for rank, hist, _, _, num_dist, hb in res:
        assert num_dist >= 2 and hb > 0
        assert len(hist) == len(want) and np.max(np.abs(hist - want)) <= HIST_TOL
    true = O.norm2(O.spgemv(h.A[0], u, b, -1.0, 1.0)) / O.norm2(b)
    assert abs(true - want[-1]) <= 1e-11
