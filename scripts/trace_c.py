"""dev: median task duration of the line-systolic solve (option trace) -> ns per step.  usage: trace_c.py <problem>"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("ILUB_PKG_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P
name = sys.argv[1]
ilupp.set_option("exact", 0)
A = {"p2d256": lambda: P.poisson2d(256), "p3d128": lambda: P.poisson3d(128), "cd3d160": lambda: P.convdiff3d(160)}[name]()
n = A.shape[0]
B = ilupp.ILU0Preconditioner(A)
x0 = torch.from_numpy(P.rhs(n, 0)).cuda(); x = x0.clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(8):
    x.copy_(x0); torch.cuda.synchronize(); e0.record(); B.apply(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
ilupp.set_option("trace", 1)
x.copy_(x0); torch.cuda.synchronize()
B.apply(x); torch.cuda.synchronize()
out = "apply %.1f us |" % np.median(ts[2:])
for which in (0, 1):
    dims, tr = B.pr.debug_sys_trace(which)
    ntasks, nper = dims[0], dims[1]
    full = np.asarray(tr).reshape(ntasks, 3, nper // 3).astype(np.float64)
    t = full[:, 0, :]
    dur = t[:, -1] - t[:, 0]
    nsteps = (nper // 3 - 1) * 8
    out += " [in-step %.0f cyc/step, waits %.0f cyc/step]" % (np.median(full[:, 1, -1]) / nsteps, np.median(full[:, 2, -1]) / nsteps)
    out += " solve %d: task %.1f us = %.0f ns/step (p10 %.0f p90 %.0f), span %.1f us |" % (
        which, np.median(dur) / 1e3, np.median(dur) / nsteps, np.percentile(dur, 10) / nsteps, np.percentile(dur, 90) / nsteps, (t.max() - t.min()) / 1e3)
print(out)
