"""Timings of the persistent asynchronous kernel under its options (factorised / explicit level-0 transfers, -read_type res,
-res_compute_type global, -async_type semi), with the per-group kernel times and CTA groups after balancing.
    python tools/async_time.py --n 256 --corrections 40 [--variants default,explicit,global,read_res,semi]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import async_multigrid_b200 as amg  # noqa: E402
from async_multigrid_b200 import hierarchy as H  # noqa: E402

VARIANTS = {
    "default": dict(factor_level0=True),
    "explicit": dict(factor_level0=False),
    "no_sellu": dict(factor_level0=True, sell_uniform=0),
    "global": dict(factor_level0=True, res_compute_type=1),
    "read_res": dict(factor_level0=True, read_type=1),
    "semi": dict(factor_level0=True, async_type=1),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--w", type=float, default=0.9)
    ap.add_argument("--corrections", type=int, default=40)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--variants", default="default,explicit,global,read_res")
    a = ap.parse_args()
    t0 = time.time()
    A = H.laplacian("7pt", a.n)
    h = H.amg_setup(A)
    b = H.rand_rhs(A.nrows)
    out = {"n": a.n, "rows": h.n, "corrections": a.corrections}
    for tag in a.variants.split(","):
        kw = dict(VARIANTS[tag])
        fact = kw["factor_level0"]
        hh = H.Hierarchy(h.A, h.P_plain)
        hh.cpts = h.cpts
        hh.build_transfers(H.MULTADD, a.w, factor_level0=fact)
        s = amg.Solver(hh, H.ASYNC_MULTADD, H.JACOBI, a.w, **kw)
        best = None
        for _ in range(a.reps + 1):          # the first solve also balances the CTA groups
            r = s.SMEM_Solve(b, 1e-9, a.corrections)
            if best is None or r["seconds"] < best["seconds"]:
                best = r
        gb = sum(H.bytes_async_chain(hh, k, True, fact) for k in range(hh.num_levels)) * a.corrections / 1e9
        out[tag] = {"seconds": best["seconds"], "relres": float(best["relres"]), "corrections": [int(x) for x in best["corrections"]],
                    "group_seconds": [round(float(x), 4) for x in s.async_group_times()], "cta_groups": [int(x) for x in s.async_groups()[0]],
                    "algorithmic_GB": round(gb, 1), "GBps": round(gb / best["seconds"])}
        s.close()
    out["host_s"] = round(time.time() - t0, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
