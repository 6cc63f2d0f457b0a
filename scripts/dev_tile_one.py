"""one grid through the tile-systolic kernel (dev tool): python scripts/dev_tile_one.py nx ny nz [exact] [key=value ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilupp_b200 as ilupp  # noqa: E402
import oracle  # noqa: E402
from ilupp_b200 import problems as P  # noqa: E402
from scripts.dev_tile_grid import grid3d  # noqa: E402

nx, ny, nz = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
exact = int(sys.argv[4]) if len(sys.argv) > 4 else 1
for kv in sys.argv[5:]:
    k, v = kv.split("=")
    ilupp.set_option(k, int(v))
ilupp.set_option("exact", exact)
A = grid3d(nx, ny, nz, 1)
n = A.shape[0]
b = P.rhs(n, 0)
B = ilupp.ILU0Preconditioner(A)
print("stats", B.stats, flush=True)
Lo, Uo = oracle.port.ilu0(A)
for tr in (False, True):
    y = b.copy()
    (B.apply_trans if tr else B.apply)(y)
    yo = oracle.port.apply_lu(Lo, Uo, b, tr)
    err = np.abs(y - yo).max() / np.abs(yo).max()
    print("tr", tr, "relerr %.3e" % err, "bitexact", np.array_equal(y, yo), flush=True)
    if err > 1e-12:
        bad = np.flatnonzero(np.abs(y - yo) > 1e-12 * np.abs(yo).max())
        print("  first bad rows", bad[:10], "count", bad.size, " (i,j,k) of first:", bad[0] % nx, (bad[0] // nx) % ny, bad[0] // (nx * ny))
