"""dev: apply time of small tile-systolic problems (per-step cost without / with hand-offs).
usage: dev_tile_time.py "nx ny nz" ["nx ny nz" ...] [key=value ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ilupp_b200 as ilupp  # noqa: E402
from ilupp_b200 import problems as P  # noqa: E402
from scripts.dev_tile_grid import grid3d  # noqa: E402

grids = [a for a in sys.argv[1:] if "=" not in a]
for kv in [a for a in sys.argv[1:] if "=" in a]:
    k, v = kv.split("=")
    ilupp.set_option(k, int(v))
ilupp.set_option("exact", 0)
for g in grids:
    nx, ny, nz = (int(t) for t in g.split())
    A = grid3d(nx, ny, nz, 1)
    n = A.shape[0]
    B = ilupp.ILU0Preconditioner(A)
    x0 = torch.from_numpy(P.rhs(n, 0)).cuda()
    x = x0.clone()
    ts = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        for i in range(8):
            x.copy_(x0)
            torch.cuda.synchronize()
            e0.record()
            B.apply(x)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = float(np.median(ts[2:]))
        steps = nx + 16
        print("%-14s mode %d: apply %.1f us = %.1f us per solve; %.0f ns per step (%d steps/solve)" % (
            g, B.stats["part_mode"], t, t / 2, t / 2 / steps * 1e3, steps), flush=True)
        if os.environ.get("TILE_TRACE"):
            ilupp.set_option("trace", 1)
            x.copy_(x0)
            B.apply(x)
            torch.cuda.synchronize()
            ilupp.set_option("trace", 0)
            dims, tr = B.pr.debug_sys_trace(0)
            ntasks, nper, nty, ntz = dims[:4]
            full = np.asarray(tr).reshape(ntz, nty, 3, nper // 3).astype(np.float64)
            dur = (full[:, :, 0, -1] - full[:, :, 0, 0]) / 1e3
            print("   tiles %dx%d: duration (us) %s" % (nty, ntz, np.round(dur.ravel()[:8], 1)))
            print("   data wait (us) %s ; gate wait (us) %s" % (np.round(full[:, :, 1, -1].ravel()[:8] / 1965, 1), np.round(full[:, :, 2, -1].ravel()[:8] / 1965, 1)))
    except Exception as e:
        print(g, "FAILED", str(e)[:100], flush=True)
