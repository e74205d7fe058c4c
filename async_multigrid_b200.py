"""Alias so the package directory `async-multigrid_b200/` (not a valid identifier) can be
imported with a normal `import async_multigrid_b200 as amg`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("async-multigrid_b200")
sys.modules[__name__] = _pkg
