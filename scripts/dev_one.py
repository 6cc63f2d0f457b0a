"""dev: time / profile the systolic apply on one problem.  usage: dev_one.py <name> [reps]
names: strip (32 lines x 4096), p2d256, p3d128, cd3d160"""
import os, sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.environ.get("ILUB_PKG_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P

def grid2d(nx, ny):
    T = lambda n: sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1])
    A = (sp.kron(sp.identity(ny), T(nx)) + sp.kron(T(ny), sp.identity(nx))).tocsr()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32); A.indptr = A.indptr.astype(np.int32)
    return A

name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ilupp.set_option("exact", 0)
if os.environ.get("ILUB_SYS_PLANE"):
    ilupp.set_option("sys_plane", int(os.environ["ILUB_SYS_PLANE"]))
A = {"strip": lambda: grid2d(4096, 32), "strip2": lambda: grid2d(2048, 64), "p2d256": lambda: P.poisson2d(256),
     "p3d128": lambda: P.poisson3d(128), "cd3d160": lambda: P.convdiff3d(160)}[name]()
n = A.shape[0]
B = ilupp.ILU0Preconditioner(A)
x0 = torch.from_numpy(P.rhs(n, 0)).cuda(); x = x0.clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(reps + 2):
    x.copy_(x0); torch.cuda.synchronize(); e0.record(); B.apply(x); e1.record(); torch.cuda.synchronize()
    if i >= 2: ts.append(e0.elapsed_time(e1))
st = B.stats
print("%s n=%d mode %s ctas %s grid %s: apply med %.1f us" % (name, n, st["part_mode"], st["nparts"], st["grid"], np.median(ts) * 1e3))
x.copy_(x0); B.apply(x); torch.cuda.synchronize()
print("checksum: sum %.15e  l2 %.15e  absmax %.15e" % (float(x.sum()), float(x.norm()), float(x.abs().max())))
