"""Print the plan of the row-partitioned asynchronous solve (amgb_dist_async_plan, host-only: no GPU needed) for one rank and
one level group -- the listing quoted in DESIGN.md section 6d.

    python tools/plan_dump.py --n 12 --ranks 2 --rank 0 --group 3 [--no-fact0] [--coarse-solve]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import async_multigrid_b200 as amg  # noqa: E402,F401
from async_multigrid_b200 import hierarchy as H, partition as PT, solver as S  # noqa: E402

NAMES = ["SPMV", "SCALE", "COPY", "ZERO", "UPDATE", "COUNT_STOP", "LOCK", "UNLOCK", "JGS", "ASYNC_GS", "PUSH", "SIGNAL", "WAIT"]
KINDS = ["F", "U", "RS", "R", "E", "T", "W", "UL", "T0", "FACC", "WS", "INVL1"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=12)
    ap.add_argument("--ranks", type=int, default=2)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--group", type=int, default=None, help="level group (default: the first replicated level)")
    ap.add_argument("--min-rows", type=int, default=40)
    ap.add_argument("--no-fact0", action="store_true")
    ap.add_argument("--coarse-solve", action="store_true")
    a = ap.parse_args()
    fact0 = not a.no_fact0
    A = H.laplacian("7pt", a.n)
    h = H.amg_setup(A)
    h.build_transfers(H.MULTADD, 0.9, factor_level0=fact0)
    shared = PT.plan_layouts(h, a.ranks, a.n * a.n, a.min_rows)
    plans = [PT.RankPlan(h, a.ranks, p, plan=shared) for p in range(a.ranks)]
    progs, slot_off, slot_group, slot_vec = S.dist_async_plan([pl.layouts for pl in plans], a.rank, H.ASYNC_MULTADD, H.JACOBI, True, fact0,
                                                              coarse_solve=a.coarse_solve)

    def vn(slot):
        if slot == -1:
            return "-"
        if slot == -2:
            return "f"
        if slot == -3:
            return "u"
        if slot <= -100:
            return "ws%d" % (-100 - slot)
        v = int(slot_vec[slot])
        return "%s%d" % (KINDS[v // 64], v % 64)

    q = shared[1] if a.group is None else a.group
    print("7-pt %d^3: %d levels, %d partitioned; rank %d of %d, group of level %d; arena %d doubles in %d slots"
          % (a.n, h.num_levels, shared[1], a.rank, a.ranks, q, slot_off[-1], len(slot_vec)))
    for op in progs[q]:
        t = NAMES[op.type]
        if op.type == 0:
            m = {0: "A", 1: "P", 2: "R", 3: "Ainv"}[op.mat_kind]
            print("%-10s %s%d%s  x=%s y=%s red=%s copy=%s  barrier=%d" % (t, m, op.mat_level, "*diag(w/d)" if op.sval else "", vn(op.slot[0]),
                                                                       vn(op.slot[1]), vn(op.slot[7]), vn(op.slot[8]), op.barrier))
        elif op.type == 10:
            print("%-10s %s[%d:+%d] -> rank %d [%d:]" % (t, vn(op.slot[0]), op.elem[0], op.count, op.dst_rank, op.elem[1]))
        elif op.type == 11:
            print("%-10s -> rank %d%s" % (t, op.dst_rank, "  (new exchange step)" if op.count else ""))
        elif op.type == 12:
            print("%-10s <- rank %d%s" % (t, op.dst_rank, "  (then group barrier)" if op.barrier else ""))
        else:
            print(t)


if __name__ == "__main__":
    main()
