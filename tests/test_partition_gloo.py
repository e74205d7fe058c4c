"""Host logic of the multi-GPU (DMEM replacement) path, on CPU: the row partition / extended numbering /
halo plan of async-multigrid_b200/partition.py executed by the numpy emulator over torch.distributed
(gloo, world_size 2 and 3) must reproduce the GLOBAL oracle cycle and solve history.  csrc/dist.cu executes
the same plan with the sm_100a kernels over NCCL."""
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


def _worker(rank, world, port, prob, n, min_rows, q, solver=2, smoother=0, w=0.9, post=1):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch.distributed as dist
    import async_multigrid_b200 as amg  # noqa: F401
    from async_multigrid_b200 import hierarchy as H, partition as PT
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A = H.laplacian(prob, n)
        h = H.amg_setup(A)
        h.build_transfers(solver, w, smooth_interp_type=smoother, num_pre=1, num_post=post)
        b = H.rand_rhs(A.nrows)
        plane = n * n if prob != "5pt" else n
        plan = PT.RankPlan(h, world, rank, plane=plane, min_rows_per_rank=min_rows)
        lay0 = plan.layouts[0]
        em = PT.DistEmulator(plan, PT.TorchComm(), w, solver, smoother, symmetric=post > 0)
        # one cycle on the right-hand side vs the global oracle
        r0 = em.new_vec(0)
        em.owned(0, r0)[:] = b[lay0.row_start:lay0.row_start + lay0.n_owned]
        c = em.owned(0, em.cycle(r0))
        pb = O.Problem(h, solver, smoother, w, num_pre=1, num_post=post)
        want = pb.cycle(b)[lay0.row_start:lay0.row_start + lay0.n_owned]
        err_cycle = float(np.max(np.abs(c - want)) / np.max(np.abs(want)))
        # whole solve: history vs the global oracle
        ncyc = 10 if solver == H.BPX else 100              # BPX alone does not converge: compare the first cycles
        u, hist = em.solve(b[lay0.row_start:lay0.row_start + lay0.n_owned], 1e-9, ncyc)
        _, want_hist, _ = pb.solve_sync(b, 1e-9, ncyc)
        ok_len = len(hist) == len(want_hist)
        err_hist = float(np.max(np.abs(hist - want_hist) / np.maximum(want_hist, 1.0))) if ok_len else 1.0
        # DMEM_AsyncSmooth in lock step = global (L1-)Jacobi; and the slot a rank's low boundary goes to in its lower
        # neighbour's vector (DistSolver.ipc_open_neighbours) is that neighbour's first ghost_hi entry
        if solver == H.MULTADD:
            x = em.async_smooth_lockstep(b[lay0.row_start:lay0.row_start + lay0.n_owned], 12)
            want_x = O.smooth("l1_jacobi" if smoother == H.L1_JACOBI else "jacobi", h.A[0], b, w, sweeps=12, zero_flag=1,
                              l1=h.l1_norms()[0])[lay0.row_start:lay0.row_start + lay0.n_owned]
            assert np.max(np.abs(x - want_x)) <= 1e-13 * np.max(np.abs(want_x))
            if rank > 0:
                nb = PT.rank_layouts(h, world, rank - 1, plan.starts, plan.num_dist, plan.halos)[0]
                assert plan.halos[0][rank - 1][0] + plan.all_counts[0][rank - 1] == nb.halo_lo + nb.n_owned
                assert nb.halo_hi == lay0.send_lo
        q.put((rank, plan.num_dist, err_cycle, err_hist, ok_len, [l.n_owned for l in plan.layouts],
               [(l.halo_lo, l.halo_hi) for l in plan.layouts]))
    finally:
        dist.destroy_process_group()


# (solver, smoother, weight, post sweeps): Multadd symmetrised / plain / L1, AFACx (src/SEQ_AMG.cpp:172-208), BPX
VARIANTS = {"multadd": (2, 0, 0.9, 1), "multadd_plain": (2, 0, 0.9, 0), "multadd_l1": (2, 6, 0.9, 1),
            "afacx": (1, 0, 0.6, 1), "afacx_l1": (1, 6, 0.9, 1), "bpx": (3, 0, 0.6, 1)}


@pytest.mark.parametrize("world,prob,n,min_rows,variant", [
    (2, "7pt", 16, 64, "multadd"), (3, "7pt", 18, 32, "multadd"), (2, "5pt", 48, 100, "multadd"),
    (2, "7pt", 16, 64, "afacx"), (3, "7pt", 18, 32, "afacx_l1"), (2, "7pt", 16, 64, "multadd_l1"),
    (2, "7pt", 16, 64, "multadd_plain"), (2, "5pt", 48, 100, "bpx")])
def test_distributed_plan_reproduces_global_cycle(world, prob, n, min_rows, variant):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, prob, n, min_rows, q) + VARIANTS[variant]) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, num_dist, err_cycle, err_hist, ok_len, owned, halos in res:
        assert num_dist >= 2, (num_dist, owned)          # at least two levels are really partitioned
        assert num_dist < len(owned)                      # and the coarse tail is replicated
        assert err_cycle <= 1e-12, err_cycle
        assert ok_len and err_hist <= 1e-10, err_hist


def test_partition_layout_properties():
    sys.path.insert(0, ROOT)
    import async_multigrid_b200 as amg  # noqa: F401
    from async_multigrid_b200 import hierarchy as H, partition as PT
    A = H.laplacian("7pt", 16)
    h = H.amg_setup(A)
    h.build_transfers(H.MULTADD, 0.9)
    for world in (1, 2, 4):
        starts, num_dist, halos = PT.plan_layouts(h, world, plane=256, min_rows_per_rank=64)
        for l, s in enumerate(starts):
            assert s[0] == 0 and s[-1] == h.n[l] and np.all(np.diff(s) >= 0)
        if world == 1:
            assert num_dist == 0
            continue
        assert np.all(starts[0] % 256 == 0)                # whole z-planes on level 0
        for rank in range(world):
            plan = PT.RankPlan(h, world, rank, plan=(starts, num_dist, halos))
            for l, lay in enumerate(plan.layouts):
                if lay.distributed:
                    # every column of the local blocks lands inside the extended range; the diagonal sits at halo_lo + i
                    a = plan.A[l]
                    assert a.ncols == lay.n_ext
                    assert np.all(a.indices[a.indptr[:-1]] == lay.halo_lo + np.arange(lay.n_owned))
                    # what I send is what my neighbour expects
                    if rank > 0:
                        assert lay.send_lo == halos[l][rank - 1][1]
                    if rank < world - 1:
                        assert lay.send_hi == halos[l][rank + 1][0]
                else:
                    assert plan.A[l].nrows == h.n[l]


@pytest.mark.parametrize("world,n", [(4, 8), (8, 8)])
def test_bench_plan_of_the_weak_series_is_consistent(tmp_path, world, n):
    """the grids of bench.py's two series (dist_bench.weak_dims: n x n x nN; the strong record's cube) cut into z-slabs for 4 and 8
    ranks: every level's rows are dealt completely, a rank's ghosts are exactly what its neighbours send, and every column of
    its row blocks lies inside its extended index space"""
    sys.path.insert(0, ROOT)
    import async_multigrid_b200 as amg  # noqa: F401
    from async_multigrid_b200 import hierarchy as H, partition as PT, dist_bench as DB
    from oracle import oracle as O
    for dims in (DB.weak_dims(n, world), (2 * n, 2 * n, 2 * n)):
        _check_plan(H, PT, dims, world)


def _check_plan(H, PT, dims, world):
    A = H.laplacian("7pt", *dims)
    h = H.amg_setup(A)
    h.build_transfers(H.MULTADD, 0.9, factor_level0=True)      # what the bench uploads
    starts, num_dist, halos = PT.plan_layouts(h, world, plane=dims[0] * dims[1], min_rows_per_rank=16)
    lays = [PT.rank_layouts(h, world, r, starts, num_dist, halos) for r in range(world)]
    assert num_dist >= 1
    for l in range(h.num_levels):
        assert sum(lays[r][l].n_owned for r in range(world)) == h.n[l]
        for r in range(world):
            lay = lays[r][l]
            if lay.distributed:
                # a rank's ghosts are exactly what its neighbours send
                assert lay.halo_lo == (lays[r - 1][l].send_hi if r > 0 else 0)
                assert lay.halo_hi == (lays[r + 1][l].send_lo if r < world - 1 else 0)
                # and every column of its row block lies inside [ghost_lo | owned | ghost_hi]
                blk = PT._block(h.A[l], lay.row_start, lay.row_start + lay.n_owned, lay.base, lay.n_ext)
                assert blk.indices.min() >= 0 and blk.indices.max() < lay.n_ext


def test_bench_plan_roundtrip_through_disk(tmp_path):
    """bench.py's multi-GPU leg hands the per-rank blocks over through files: what a rank loads must be what
    RankPlan builds in memory"""
    sys.path.insert(0, ROOT)
    import types
    import async_multigrid_b200 as amg  # noqa: F401
    from async_multigrid_b200 import hierarchy as H, partition as PT, dist_bench as DB
    args = types.SimpleNamespace(n=12, theta=0.25, smooth_weight=0.9, num_post=1, min_rows_per_rank=64)
    world = 2
    d = str(tmp_path / "plan")
    dims = DB.weak_dims(12, world)
    assert dims == (12, 12, 24) and DB.weak_dims(256, 8) == (256, 256, 2048)
    DB._build_and_scatter(args, world, d, dims)
    A = H.laplacian("7pt", *dims)
    h = H.amg_setup(A)
    h.build_transfers(H.MULTADD, 0.9, factor_level0=True)      # the bench's multi-GPU leg uploads plain P_0 / R_0
    b = H.rand_rhs(A.nrows)
    for rank in range(world):
        got = DB._PlanFromDisk(d, rank, world)
        want = PT.RankPlan(h, world, rank, plane=144, min_rows_per_rank=64)
        assert got.num_dist == want.num_dist and got.num_levels == want.num_levels
        for l in range(want.num_levels):
            a, bb = got.layouts[l], want.layouts[l]
            assert (a.n_global, a.row_start, a.n_owned, a.halo_lo, a.halo_hi, a.distributed, a.send_lo, a.send_hi) == \
                   (bb.n_global, bb.row_start, bb.n_owned, bb.halo_lo, bb.halo_hi, bb.distributed, bb.send_lo, bb.send_hi)
            for mg, mw in ((got.A[l], want.A[l]),) + (((got.P[l], want.P[l]), (got.R[l], want.R[l])) if l < want.num_levels - 1 else ()):
                assert mg.shape == mw.shape and np.array_equal(mg.indptr, mw.indptr)
                assert np.array_equal(mg.indices, mw.indices) and np.array_equal(mg.data, mw.data)
        l0 = want.layouts[0]
        assert np.array_equal(got.b, b[l0.row_start:l0.row_start + l0.n_owned])


def test_bench_calibration_of_the_asynchronous_correction_count():
    """dist_bench._calibrate_corrections (bench.py --gpus N, asynchronous leg): from the synchronous cycle count down in steps of
    2 while the solve converges with a margin, else up in steps of 4, at most five tries"""
    sys.path.insert(0, ROOT)
    import async_multigrid_b200 as amg  # noqa: F401
    from async_multigrid_b200 import dist_bench as DB
    calls = []

    def one(K):
        calls.append(K)
        return [K] * 3, 0.6 ** K, 0.001 * K

    K, tried = DB._calibrate_corrections(one, 46, 1e-9)
    assert K == 42 and [t["corrections"] for t in tried] == [46, 44, 42, 40]          # 0.6^40 = 1.3e-9 misses, 0.6^42 = 4.8e-10 < 0.7e-9
    K, tried = DB._calibrate_corrections(one, 30, 1e-9)
    assert K == 42 and [t["corrections"] for t in tried] == [30, 34, 38, 42]
    K, tried = DB._calibrate_corrections(one, 20, 1e-9)
    assert K == 36 and len(tried) == 5 and tried[-1]["relres"] > 1e-9                   # gives up: the record then says converged = false
    K, tried = DB._calibrate_corrections(one, 60, 1e-9)
    assert K == 52 and len(tried) == 5                                                  # never more than five solves
    assert all(t["seconds"] == 0.001 * t["corrections"] for t in tried)
