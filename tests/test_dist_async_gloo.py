"""The row-partitioned asynchronous solve's plans (csrc/dist_async.cu dist_async_plan) executed by SEPARATE PROCESSES over
torch.distributed (gloo, world_size 2, 3 and 4): every process interprets only its own rank's programs on its own arena;
AOP_PUSH becomes a message to the destination rank, AOP_SIGNAL a message carrying the exchange-step number, AOP_WAIT drains
the source rank's messages until the step has arrived.  That is the device protocol with the NVLink stores replaced by
messages: a rank sees a peer's boundary values only through the pushes the plan contains, and can only proceed past a wait
when the peer really signalled -- a push or a signal missing from the plan shows up as a wrong result or a hang here.  With the
groups taking turns the result must equal the single-GPU programs on the unpartitioned hierarchy (tests/async_emulator.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, min_rows, fact0, K, q_out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch
    import torch.distributed as dist
    import async_multigrid_b200 as amg  # noqa: F401
    from async_multigrid_b200 import hierarchy as H
    from dist_async_emulator import DistAsyncEmulator, PUSH, SIGNAL, WAIT, X, Y
    from test_dist_async_plan import _problem
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, b, plane = _problem("7pt", n, fact0=fact0)

        class OneRank(DistAsyncEmulator):
            """this process is rank `rank`: the other ranks' arenas are never touched"""
            pending = []

            def exec_op(self, p, q, i):
                assert p == rank
                op = self.progs[p][q][i]
                if op.type == PUSH:
                    if op.count > 0:
                        hd = torch.tensor([0, op.slot[Y], op.elem[Y], op.count], dtype=torch.int64)
                        data = torch.from_numpy(self.view(p, op.slot[X], op.elem[X], op.count).copy())
                        self.pending += [(hd, dist.isend(hd, op.dst_rank)), (data, dist.isend(data, op.dst_rank))]
                        self.pushed += op.count
                elif op.type == SIGNAL:
                    if op.count:
                        self.seq[p][q] += 1
                    if op.dst_rank >= 0:
                        hd = torch.tensor([1, q, self.seq[p][q], 0], dtype=torch.int64)
                        self.pending.append((hd, dist.isend(hd, op.dst_rank)))
                elif op.type == WAIT:
                    src = op.dst_rank
                    while src >= 0 and self.flag[p][q, src] < self.seq[p][q]:
                        hd = torch.zeros(4, dtype=torch.int64)
                        dist.recv(hd, src)
                        if hd[0] == 0:
                            data = torch.zeros(int(hd[3]), dtype=torch.float64)
                            dist.recv(data, src)
                            self.view(p, int(hd[1]), int(hd[2]), int(hd[3]))[:] = data.numpy()
                        else:
                            self.flag[p][int(hd[1]), src] = int(hd[2])
                else:
                    super().exec_op(p, q, i)

        em = OneRank(h, world, b, H.ASYNC_MULTADD, H.JACOBI, 0.9, factor_level0=fact0, plane=plane, min_rows_per_rank=min_rows)
        for _ in range(K):
            for q in range(em.L):
                for i in range(len(em.progs[rank][q])):
                    em.exec_op(rank, q, i)
        for _, w in em.pending:
            w.wait()
        l0 = em.plans[rank].layouts[0]
        mine = em.u[rank][l0.halo_lo:l0.halo_lo + l0.n_owned].copy()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        q_out.put((rank, np.concatenate(parts), em.num_dist, int(em.pushed), em.count[rank]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,min_rows,fact0", [(2, 12, 40, True), (3, 12, 40, False), (2, 16, 64, True), (4, 16, 40, True)])
def test_plans_run_by_separate_processes_equal_the_single_gpu_programs(world, n, min_rows, fact0):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import async_multigrid_b200 as amg  # noqa: F401
    from async_multigrid_b200 import hierarchy as H, solver as S
    from async_emulator import Emulator
    from test_dist_async_plan import _problem
    K = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, min_rows, fact0, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=240) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    h, b, _ = _problem("7pt", n, fact0=fact0)
    want = Emulator(h, S.async_program(h.num_levels, H.ASYNC_MULTADD, H.JACOBI, symmetric=True, factor_level0=fact0), b, H.JACOBI, 0.9).run(K)
    for rank, u, num_dist, pushed, count in res:
        assert num_dist >= 1 and pushed > 0
        assert count == [K] * h.num_levels
        assert np.max(np.abs(u - want)) <= 1e-12 * np.max(np.abs(want))
