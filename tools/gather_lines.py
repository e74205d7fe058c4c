"""CPU model of the x-gather cost that bounds the Galerkin / transfer operators (profiles/README.md section 3): distinct
128-byte lines of x touched per warp-wide gather when 32 consecutive rows sit in the 32 lanes (sliced ELL), in the natural
ordering and in a tiled ordering of the unknowns.  python tools/gather_lines.py --n 64 --tile 4"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from async_multigrid_b200 import hierarchy as H  # noqa: E402


def lines_per_gather(m):
    """(mean distinct 128-B lines per warp gather, mean active lanes per gather) over slices of 32 consecutive rows"""
    n = m.nrows
    lens = np.diff(m.indptr)
    tot_lines, tot_gathers, tot_lanes = 0, 0, 0
    for s in range(0, n, 32):
        rows = np.arange(s, min(s + 32, n))
        width = int(lens[rows].max()) if rows.size else 0
        for j in range(width):
            act = rows[lens[rows] > j]
            cols = m.indices[m.indptr[act] + j]
            tot_lines += np.unique(cols // 16).size
            tot_lanes += act.size
            tot_gathers += 1
    return tot_lines / max(tot_gathers, 1), tot_lanes / max(tot_gathers, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=48)
    ap.add_argument("--tile", type=int, default=4)
    ap.add_argument("--sample", type=int, default=200000, help="rows per matrix (leading slices)")
    a = ap.parse_args()
    A = H.laplacian("7pt", a.n)
    h = H.amg_setup(A)
    h2, _ = H.reorder_hierarchy(h, H.tiled_permutation(a.n, a.n, a.n, a.tile))
    for hh in (h, h2):
        hh.build_transfers(H.MULTADD, 0.9)
    print("n=%d tile=%d levels %s" % (a.n, a.tile, h.n))
    for name, get in (("A0", lambda x: x.A[0]), ("Rbar0", lambda x: x.R[0]), ("Pbar0", lambda x: x.P[0]), ("P0 plain", lambda x: x.P_plain[0]),
                      ("A1", lambda x: x.A[1]), ("Pbar1", lambda x: x.P[1]), ("Rbar1", lambda x: x.R[1])):
        out = []
        for hh in (h, h2):
            m = get(hh)
            # sorted columns inside a row, as the SELL builder stores them
            sm = H.CSR.from_scipy(m.to_scipy().copy())
            if sm.nrows > a.sample:
                k = a.sample // 32 * 32
                mid = (sm.nrows // 2) // 32 * 32
                lo = max(0, mid - k // 2)
                sub = H.CSR(k, sm.ncols, sm.indptr[lo:lo + k + 1] - sm.indptr[lo], sm.indices[sm.indptr[lo]:sm.indptr[lo + k]],
                            sm.data[sm.indptr[lo]:sm.indptr[lo + k]])
                sm = sub
            out.append(lines_per_gather(sm))
        print("%-9s natural: %5.2f lines / gather (%4.1f lanes)   tiled: %5.2f lines / gather (%4.1f lanes)" %
              (name, out[0][0], out[0][1], out[1][0], out[1][1]))


if __name__ == "__main__":
    main()
