"""A/B of the persistent asynchronous kernel: explicit level-0 products (k_async_amg, the measured round-1 kernel) against the
factorised level-0 transfers (k_async_amg_fact0, experimental) and against the same kernel with non-inlined SpMV calls
(k_async_amg_ni, AMGB_ASYNC_NOINLINE=1, experimental: code size).  python tools/async_fact0_time.py --n 256 --corrections 40"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import async_multigrid_b200 as amg  # noqa: E402
from async_multigrid_b200 import hierarchy as H  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--w", type=float, default=0.9)
    ap.add_argument("--corrections", type=int, default=40)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    t0 = time.time()
    A = H.laplacian("7pt", a.n)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    out = {"n": a.n, "rows": h.n, "corrections": a.corrections}
    for tag, fact in (("explicit", False), ("explicit_noinline", False), ("factorised", True)):
        os.environ["AMGB_ASYNC_NOINLINE"] = "1" if tag == "explicit_noinline" else "0"
        hh = H.Hierarchy(h.A, h.P_plain)
        hh.cpts = h.cpts
        hh.build_transfers(H.MULTADD, a.w, factor_level0=fact)
        s = amg.Solver(hh, H.ASYNC_MULTADD, H.JACOBI, a.w, factor_level0=fact)
        best = None
        for _ in range(a.reps):
            r = s.SMEM_Solve(b, 1e-9, a.corrections)
            if best is None or r["seconds"] < best["seconds"]:
                best = r
        out[tag] = {"seconds": best["seconds"], "relres": float(best["relres"]), "corrections": [int(x) for x in best["corrections"]]}
        s.close()
    out["host_s"] = round(time.time() - t0, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
