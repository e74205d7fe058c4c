"""B200-native solve phase for async-multigrid's additive AMG cycles (Multadd, AFACx, BPX).

Importable as ``importlib.import_module("async-multigrid_b200")`` or through the
``async_multigrid_b200`` alias module at the repository root.

  hierarchy  host-side input provider (problem matrices, classical AMG hierarchy, smoothed
             transfers, thread/CTA-group work model)           -- CPU, not the accelerated path
  solver     ctypes mirror of the reference's solve-phase interface over the C ABI
             (include/amg_b200.h -> libamg_b200.so, hand-written sm_100a kernels)
  build      in-tree nvcc / g++ builds
"""
from . import build, hierarchy, partition, solver  # noqa: F401
from .solver import Solver, DistSolver, ExtendedExplicitSolver, AmgError  # noqa: F401
