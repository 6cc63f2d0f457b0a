"""dev: general (non-stencil) solve kernels on a randomly permuted 128^3 Poisson matrix and on the lexicographic one.
usage: dev_general.py [N]"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("ILUB_PKG_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
A = P.poisson3d(N)
n = A.shape[0]
perm = np.random.default_rng(1).permutation(n)
Ap = A[perm][:, perm].tocsr(); Ap.sort_indices()
b = P.rhs(n, 0)
x0 = torch.from_numpy(b).cuda()
ilupp.set_option("exact", 0)
ref = {}
for name, M in (("permuted", Ap), ("lexicographic", A)):
    for algo in (0, 1, 2, 4):
        if name == "lexicographic" and algo == 0:
            continue
        ilupp.set_option("solve_algo", algo)
        try:
            B = ilupp.IChol0Preconditioner(M)
            x = x0.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for i in range(8):
                x.copy_(x0); torch.cuda.synchronize(); e0.record(); B.apply(x); e1.record(); torch.cuda.synchronize()
                if i >= 2: ts.append(e0.elapsed_time(e1))
            y = x.cpu().numpy()
            ref.setdefault(name, y)
            by = 12 * 2 * B.total_nnz + 8 * (n + 1) + 32 * n
            st = B.stats
            print("%-14s algo %d: apply %.3f ms = %.0f GB/s | mode %s levels %s | sym %.1f ms | relerr vs first %.1e" % (
                name, algo, np.median(ts), by / np.median(ts) / 1e6, st["part_mode"], st["levels"][:2], B.timings_ms[0],
                np.abs(y - ref[name]).max() / np.abs(ref[name]).max()), flush=True)
            del B
        except Exception as e:
            print(name, algo, "FAILED", str(e)[:200])
ilupp.set_option("solve_algo", 0)
