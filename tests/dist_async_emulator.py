"""CPU interpreter of the ROW-PARTITIONED asynchronous solve's plans (amgb_dist_async_plan, csrc/dist_async.cu) -- test
infrastructure.  Every rank's persistent kernel executes the operation lists this module reads, on the rank's row blocks
(partition.RankPlan), with vectors addressed as (slot, element) of a per-rank arena and AOP_PUSH operations that store
boundary / owned entries into a PEER's arena.  Here all ranks live in one process; two schedulers:

  * lock step: the groups take turns and, inside a group's iteration, every rank executes operation i before any rank
    executes operation i + 1 -- the result must equal the single-GPU programs interpreted on the unpartitioned hierarchy
    (tests/async_emulator.py) to rounding;
  * random: (rank, group) pairs advance by random numbers of operations in random order; a pair standing at an AOP_WAIT
    whose flag has not arrived cannot advance.  The exchange steps (PUSH .. SIGNAL .. WAIT) keep the ranks of ONE group in
    step; the groups drift apart freely, as on the device.  With `honour_waits=False` the waits are ignored: that is the
    design without flags, kept to show what the flags are for (stale intermediate ghosts cost orders of magnitude)."""
import numpy as np

from async_multigrid_b200 import hierarchy as H
from async_multigrid_b200 import partition as PT
from async_multigrid_b200 import solver as S

SPMV, SCALE, COPY, ZERO, UPDATE, COUNT_STOP, LOCK, UNLOCK, JGS, ASYNC_GS, PUSH, SIGNAL, WAIT = range(13)
X, Y, B, C_, RS, B2, XS, RED, RED_COPY, ACC = range(10)
V_R, V_UL = 3, 7


class DistAsyncEmulator:
    def __init__(self, h, nranks, f, solver, smoother, w, symmetric=True, factor_level0=False, plane=None, min_rows_per_rank=64,
                 fine_sweeps=1, coarse_sweeps=1, coarse_solve=False):
        import scipy.sparse as sp
        self.h, self.P = h, nranks
        self.L = h.num_levels
        shared = PT.plan_layouts(h, nranks, plane, min_rows_per_rank)
        self.plans = [PT.RankPlan(h, nranks, p, plan=shared) for p in range(nranks)]
        self.num_dist = shared[1]
        layouts = [pl.layouts for pl in self.plans]
        self.progs, self.slot_off, self.slot_group, self.slot_vec = [], None, None, None
        for p in range(nranks):
            pr, so, sg, sv = S.dist_async_plan(layouts, p, solver, smoother, symmetric, factor_level0, fine_sweeps, coarse_sweeps, coarse_solve)
            if self.slot_off is None:
                self.slot_off, self.slot_group, self.slot_vec = so, sg, sv
            else:       # the slot table must be the same on every rank: peers are addressed through it
                assert np.array_equal(so, self.slot_off) and np.array_equal(sg, self.slot_group) and np.array_equal(sv, self.slot_vec)
            self.progs.append(pr)
        for q in range(self.L):
            assert len(set(len(self.progs[p][q]) for p in range(nranks))) == 1, "programs of one group differ in length between ranks"
        self.arena = [np.zeros(int(self.slot_off[-1])) for _ in range(nranks)]
        # global scale vectors -> every rank's level layout (ghosts included)
        gws = [w / a.to_scipy().diagonal() for a in h.A]
        if smoother == H.L1_JACOBI:
            gws = [1.0 / x for x in h.l1_norms()]
        self.A, self.Pm, self.R, self.Asv, self.ws = [], [], [], [], []
        for p, pl in enumerate(self.plans):
            self.A.append([m.to_scipy() for m in pl.A])
            self.Pm.append([m.to_scipy() for m in pl.P])
            self.R.append([m.to_scipy() for m in pl.R])
            wsp = []
            for l, lay in enumerate(pl.layouts):
                wsp.append(gws[l][lay.base:lay.base + lay.n_ext].copy())
            self.ws.append(wsp)
            self.Asv.append([a @ sp.diags(wsp[l]) for l, a in enumerate(self.A[p])])
        self.Ainv = sp.csr_matrix(np.linalg.inv(h.A[-1].to_scipy().toarray()))
        self.fg = np.asarray(f, dtype=np.float64)
        self.Ag = h.A[0].to_scipy()
        self.f = []
        self.u = []
        r0 = self.fg.copy()                  # u = 0
        self.r0_norm = np.linalg.norm(r0)
        for p, pl in enumerate(self.plans):
            l0 = pl.layouts[0]
            self.f.append(self.fg[l0.row_start:l0.row_start + l0.n_owned].copy())
            self.u.append(np.zeros(l0.n_ext))
            for s in range(len(self.slot_vec)):
                if self.slot_vec[s] == V_R * 64:      # every group starts from r0 with its ghosts (and from u = 0)
                    o = int(self.slot_off[s])
                    self.arena[p][o:o + l0.n_ext] = r0[l0.base:l0.base + l0.n_ext]
        self.count = [[0] * self.L for _ in range(nranks)]
        self.pushed = 0
        self.seq = [[0] * self.L for _ in range(nranks)]                       # exchange steps of (rank, group) so far
        self.flag = [np.zeros((self.L, nranks), dtype=np.int64) for _ in range(nranks)]      # flag[p][q, src]
        self.honour_waits = True

    # ---- operands
    def view(self, p, slot, elem, n):
        elem = int(elem)
        if slot >= 0:
            o = int(self.slot_off[slot]) + elem
            assert o + n <= int(self.slot_off[slot + 1]), "operand runs past its slot"
            return self.arena[p][o:o + n]
        if slot == -2:
            assert elem + n <= len(self.f[p])
            return self.f[p][elem:elem + n]
        if slot == -3:
            assert elem + n <= len(self.u[p])
            return self.u[p][elem:elem + n]
        if slot <= -100:
            v = self.ws[p][-100 - slot]
            assert elem + n <= len(v)
            return v[elem:elem + n]
        raise AssertionError("no such operand")

    def _store(self, p, op, n, t):
        if op.slot[RED] != -1:
            r = self.view(p, op.slot[RED], op.elem[RED], n)
            r += op.red_scale * t
            if op.slot[RED_COPY] != -1:
                self.view(p, op.slot[RED_COPY], op.elem[RED_COPY], n)[:] = r
        if op.slot[ACC] != -1:
            self.view(p, op.slot[ACC], op.elem[ACC], n)[:] += t
        if op.slot[Y] != -1:
            self.view(p, op.slot[Y], op.elem[Y], n)[:] = t

    def exec_op(self, p, q, i):
        op = self.progs[p][q][i]
        if op.type == SPMV:
            if op.mat_kind == 3:             # coarse_solve: the dense inverse of the (replicated) coarsest operator
                M = self.Ainv
            else:
                M = (self.A, self.Pm, self.R)[op.mat_kind][p][op.mat_level]
            if op.sval:
                assert op.mat_kind == 0
                M = self.Asv[p][op.mat_level]
            n, nc = M.shape
            t = op.alpha * (M @ self.view(p, op.slot[X], op.elem[X], nc))
            if op.slot[B] != -1:
                t = t + op.beta * self.view(p, op.slot[B], op.elem[B], n)
            if op.slot[B2] != -1:
                t = t + op.beta2 * self.view(p, op.slot[B2], op.elem[B2], n)
            if op.slot[RS] != -1:
                t = t * self.view(p, op.slot[RS], op.elem[RS], n)
            if op.slot[C_] != -1:
                t = t + op.gamma * self.view(p, op.slot[C_], op.elem[C_], n)
            if op.slot[XS] != -1:
                t = t + op.xself * self.view(p, op.slot[XS], op.elem[XS], n)
            self._store(p, op, n, t)
        elif op.type == SCALE:
            n = self.A[p][op.level].shape[0]
            t = self.view(p, op.slot[RS], op.elem[RS], n) * self.view(p, op.slot[X], op.elem[X], n)
            self._store(p, op, n, t)
        elif op.type == COPY:
            n = self.A[p][op.level].shape[0]
            self.view(p, op.slot[Y], op.elem[Y], n)[:] = self.view(p, op.slot[X], op.elem[X], n)
        elif op.type == ZERO:
            n = self.A[p][op.level].shape[0]
            self.view(p, op.slot[Y], op.elem[Y], n)[:] = 0.0
        elif op.type == UPDATE:
            n = self.A[p][0].shape[0]
            e = self.view(p, op.slot[X], op.elem[X], n)
            if op.slot[ACC] != -1:
                self.view(p, op.slot[ACC], op.elem[ACC], n)[:] += e
            if op.slot[RED] != -1:
                r = self.view(p, op.slot[RED], op.elem[RED], n)
                r += op.red_scale * e
                if op.slot[RED_COPY] != -1:
                    self.view(p, op.slot[RED_COPY], op.elem[RED_COPY], n)[:] = r
        elif op.type == PUSH:
            if op.count > 0:
                assert 0 <= op.dst_rank < self.P and op.dst_rank != p
                self.view(op.dst_rank, op.slot[Y], op.elem[Y], op.count)[:] = self.view(p, op.slot[X], op.elem[X], op.count)
                self.pushed += op.count
        elif op.type == SIGNAL:
            if op.count:
                self.seq[p][q] += 1
            if op.dst_rank >= 0:
                assert op.dst_rank != p
                self.flag[op.dst_rank][q, p] = self.seq[p][q]
        elif op.type == WAIT:
            if op.dst_rank >= 0 and self.honour_waits:
                assert self.flag[p][q, op.dst_rank] >= self.seq[p][q], "executed a wait whose flag has not arrived"
        elif op.type == COUNT_STOP:
            self.count[p][q] += 1
        else:
            raise NotImplementedError("operation %d is not part of the row-partitioned programs" % op.type)

    # ---- schedulers
    def run_lockstep(self, num_cycles):
        for _ in range(num_cycles):
            for q in range(self.L):
                for i in range(len(self.progs[0][q])):
                    for p in range(self.P):
                        self.exec_op(p, q, i)
        return self.solution()

    def blocked(self, p, q, i):
        op = self.progs[p][q][i]
        return (self.honour_waits and op.type == WAIT and op.dst_rank >= 0 and
                self.flag[p][q, op.dst_rank] < self.seq[p][q])

    def run_random(self, num_cycles, seed=0, max_burst=6, honour_waits=True):
        """every (rank, group) advances by bursts of 1..max_burst operations in random order until each has counted
        num_cycles corrections"""
        self.honour_waits = honour_waits
        rng = np.random.default_rng(seed)
        pc = {(p, q): 0 for p in range(self.P) for q in range(self.L)}
        live = [k for k in pc if len(self.progs[k[0]][k[1]]) > 0]
        while live:
            assert any(not self.blocked(p, q, pc[(p, q)]) for p, q in live), "deadlock: every live group waits for a peer"
            p, q = live[rng.integers(len(live))]
            prog = self.progs[p][q]
            for _ in range(int(rng.integers(1, max_burst + 1))):
                if self.blocked(p, q, pc[(p, q)]):
                    break
                self.exec_op(p, q, pc[(p, q)])
                pc[(p, q)] = (pc[(p, q)] + 1) % len(prog)
                if self.count[p][q] >= num_cycles and prog[(pc[(p, q)] - 1) % len(prog)].type == COUNT_STOP:
                    live.remove((p, q))
                    break
        return self.solution()

    def solution(self):
        parts = []
        for p, pl in enumerate(self.plans):
            l0 = pl.layouts[0]
            parts.append(self.u[p][l0.halo_lo if l0.distributed else 0:][:l0.n_owned])
        return np.concatenate(parts)

    def relres(self):
        return np.linalg.norm(self.fg - self.Ag @ self.solution()) / self.r0_norm
