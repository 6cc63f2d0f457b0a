"""dev: time the batched multi-RHS apply (BASELINE config 5 shape).  usage: dev_batched.py <nsys> <nrhs> [key=value ...] [reps=N]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P

nsys, nrhs = int(sys.argv[1]), int(sys.argv[2])
reps = 5
ilupp.set_option("exact", 0)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    if k == "reps":
        reps = int(v)
    else:
        ilupp.set_option(k, int(v))
pat, vals = P.batched_poisson2d(nsys, 128)
n1 = pat.shape[0]
B = ilupp.ILU0PreconditionerBatched(pat, vals)
X = torch.randn((nrhs, nsys * n1), dtype=torch.float64, device="cuda").t()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(reps + 2):
    torch.cuda.synchronize(); e0.record(); B.apply(X); e1.record(); torch.cuda.synchronize()
    if i >= 2: ts.append(e0.elapsed_time(e1))
by = nsys * (12 * 97792 + 8 * (n1 + 1)) + 32.0 * nsys * n1 * nrhs
print("batched nsys=%d nrhs=%d %s: apply med %.3f ms min %.3f | %.2f M solves/s | %.0f GB/s algorithmic" % (
    nsys, nrhs, " ".join(sys.argv[3:]), np.median(ts), min(ts), nsys * nrhs / np.median(ts) / 1e3, by / np.median(ts) / 1e6))
