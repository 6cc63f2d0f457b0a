"""dev: block-sequential solve with several right-hand sides (one CTA per column) on the irregular matrix.
usage: dev_block_multi.py [n]   (host arrays: the times include the PCIe copies of the columns)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
A = P.irregular_powerlaw(n=n)
B = ilupp.ILU0Preconditioner(A, options={"exact": 0})
print("part_mode", B.stats["part_mode"], flush=True)
for k in (1, 4, 16, 64):
    X = np.asfortranarray(P.rhs(n, 1, k=k))
    ts = []
    for i in range(3):
        Y = X.copy(order="F")
        t0 = time.perf_counter(); B.pr.apply_multi(Y, False); ts.append((time.perf_counter() - t0) * 1e3)
    print("nrhs %2d: %.1f ms = %.2f ms per right-hand side" % (k, min(ts), min(ts) / k), flush=True)
