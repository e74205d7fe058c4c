"""CPU interpreter of the persistent asynchronous kernel's per-group PROGRAMS (amgb_async_program, csrc/async.cu
async_build_program) -- test infrastructure.  The device kernel (k_async_amg) executes exactly these operation lists, one
list per level group, with group barriers in between; here a group's iteration runs atomically and the groups take turns,
which is one legal interleaving of the asynchronous method.  Gauss-Seidel-type smoother operations are not emulated (the
GPU suite covers them)."""
import numpy as np

from async_multigrid_b200 import hierarchy as H

SPMV, SCALE, COPY, ZERO, UPDATE, COUNT_STOP, LOCK, UNLOCK, JGS, ASYNC_GS = range(10)
(V_F, V_U, V_RS, V_R, V_E, V_T, V_W, V_UL, V_T0, V_FACC, V_WS, V_INVL1) = range(12)


class Emulator:
    def __init__(self, h, progs, f, smoother, w, u0=None, ctas_per_group=3, first_group=0):
        self.h, self.progs = h, progs
        self.L = h.num_levels
        self.A = [m.to_scipy() for m in h.A]
        self.P = [m.to_scipy() for m in h.P]
        self.R = [m.to_scipy() for m in h.R]
        self.ws = [w / a.diagonal() for a in h.A]
        self.inv_l1 = [1.0 / x for x in h.l1_norms()]
        rs = self.inv_l1 if smoother == H.L1_JACOBI else self.ws
        import scipy.sparse as sp
        self.Asv = [a @ sp.diags(rs[l]) for l, a in enumerate(self.A)]
        n0 = h.n[0]
        self.f = np.asarray(f, dtype=np.float64)
        self.u = np.zeros(n0) if u0 is None else np.array(u0, dtype=np.float64)
        r0 = self.f - self.A[0] @ self.u
        self.r0_norm = np.linalg.norm(r0)
        self.rs_shared = r0.copy()
        self.first = first_group
        self.g = []
        for q in range(self.L):
            d = {}
            d[(V_R, 0)] = r0.copy()
            d[(V_UL, 0)] = self.u.copy()
            d[(V_FACC, 0)] = np.zeros(n0)
            self.g.append(d)
        self.count = [0] * self.L
        # fake CTA layout for the CTA-slice operations: `ctas_per_group` CTAs per group (one for the idle coarsest group)
        self.cta_begin = [0] * (self.L + 1)
        for q in range(self.L):
            n = 0 if q < first_group else (1 if q == self.L - 1 and self.L > 1 else ctas_per_group)
            self.cta_begin[q + 1] = self.cta_begin[q] + n
        self.grid = self.cta_begin[self.L - 1] if self.L > 1 else self.cta_begin[-1]    # (the idle coarsest group owns no rows)

    def vec(self, q, vid, level_rows=None):
        if vid < 0:
            return None
        kind, l = vid // 64, vid % 64
        if kind == V_F:
            return self.f
        if kind == V_U:
            return self.u
        if kind == V_RS:
            return self.rs_shared
        if kind == V_WS:
            return self.ws[l]
        if kind == V_INVL1:
            return self.inv_l1[l]
        key = (kind, l)
        if key not in self.g[q]:
            lv = 0 if kind in (V_UL, V_T0, V_FACC) else l
            self.g[q][key] = np.zeros(self.h.n[lv])
        return self.g[q][key]

    def _rows(self, q, op):
        """row sets an operation covers: everything, or the slices of the group's CTAs"""
        n = self.h.n[0]
        if not op.range:
            return [slice(None)]
        out = []
        for cta in range(self.cta_begin[q], self.cta_begin[q + 1]):
            out.append(slice(n * cta // self.grid, n * (cta + 1) // self.grid))
        return out

    def run_group_iteration(self, q):
        for op in self.progs[q]:
            if op.type == SPMV:
                if op.mat_kind == 3:         # coarse_solve: the dense inverse of the coarsest operator
                    import scipy.sparse as sp
                    M = sp.csr_matrix(np.linalg.inv(self.A[op.mat_level].toarray()))
                else:
                    M = (self.A, self.P, self.R)[op.mat_kind][op.mat_level]
                if op.sval:
                    assert op.mat_kind == 0
                    M = self.Asv[op.mat_level]
                x = self.vec(q, op.x)
                for rows in self._rows(q, op):
                    t = op.alpha * (M[rows] @ x)
                    if op.b >= 0:
                        t = t + op.beta * self.vec(q, op.b)[rows]
                    if op.b2 >= 0:
                        t = t + op.beta2 * self.vec(q, op.b2)[rows]
                    if op.rs >= 0:
                        t = t * self.vec(q, op.rs)[rows]
                    if op.c >= 0:
                        t = t + op.gamma * self.vec(q, op.c)[rows]
                    if op.xs >= 0:
                        t = t + op.xself * self.vec(q, op.xs)[rows]
                    self._store(q, op, rows, t)
            elif op.type == SCALE:
                for rows in self._rows(q, op):
                    t = self.vec(q, op.rs)[rows] * self.vec(q, op.x)[rows]
                    self._store(q, op, rows, t)
            elif op.type == COPY:
                self.vec(q, op.y)[:] = self.vec(q, op.x)
            elif op.type == ZERO:
                self.vec(q, op.y)[:] = 0.0
            elif op.type == UPDATE:
                e = self.vec(q, op.x)
                if op.acc >= 0:
                    self.vec(q, op.acc)[:] += e
                if op.red >= 0:
                    self.vec(q, op.red)[:] += op.red_scale * e
                    if op.red_copy >= 0:
                        self.vec(q, op.red_copy)[:] = self.vec(q, op.red)
            elif op.type == COUNT_STOP:
                self.count[q] += 1
            elif op.type in (LOCK, UNLOCK):
                pass
            else:
                raise NotImplementedError("smoother operation %d is not emulated" % op.type)

    def _store(self, q, op, rows, t):
        if op.red >= 0:
            self.vec(q, op.red)[rows] += op.red_scale * t
            if op.red_copy >= 0:
                self.vec(q, op.red_copy)[rows] = self.vec(q, op.red)[rows]
        if op.acc >= 0:
            self.vec(q, op.acc)[rows] += t
        if op.y >= 0:
            self.vec(q, op.y)[rows] = t

    def run(self, num_cycles, read_res=False):
        """every group performs num_cycles iterations, the groups taking turns (LOCAL stop rule)"""
        for _ in range(num_cycles):
            for q in range(self.first, self.L):
                self.run_group_iteration(q)
        if read_res:
            for q in range(self.first, self.L):
                self.u += self.vec(q, V_FACC * 64)
        return self.u

    def relres(self):
        return np.linalg.norm(self.f - self.A[0] @ self.u) / self.r0_norm
