#!/usr/bin/env python
"""How evenly do the level groups of the persistent asynchronous kernel progress?  GLOBAL stop rule: all groups
stop when the slowest has done `cycles` corrections; the per-level counts then expose the imbalance."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import async_multigrid_b200 as amg
from async_multigrid_b200 import hierarchy as H

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 20
A = H.laplacian("7pt", n)
h = H.amg_setup(A)
h.build_transfers(H.MULTADD, 0.9)
b = H.rand_rhs(A.nrows)
s = amg.Solver(h, H.ASYNC_MULTADD, H.JACOBI, 0.9)
cb, grid = s.async_groups()
work, frac = H.compute_work(h, H.MULTADD)
s.set_rhs(b)
for rule in (amg.solver.CONVERGE_GLOBAL, amg.solver.CONVERGE_LOCAL):
    s.set_solution(None)
    corr, rel, secs = s.solve_async(cycles, rule)
    print("rule", rule, "grid", grid, "ctas/level", list(np.diff(cb)), "corrections", list(corr), "relres %.2e" % rel, "%.3fs" % secs)
print("work fractions", [round(f, 3) for f in frac])
