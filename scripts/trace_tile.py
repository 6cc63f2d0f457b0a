"""dev: stage-level timeline of the tile-systolic solve (option trace).  usage: trace_tile.py <problem> [key=value ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ilupp_b200 as ilupp  # noqa: E402
from ilupp_b200 import problems as P  # noqa: E402

name = sys.argv[1]
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ilupp.set_option(k, int(v))
ilupp.set_option("exact", 0)
A = {"p3d128": lambda: P.poisson3d(128), "p3d64": lambda: P.poisson3d(64), "cd3d160": lambda: P.convdiff3d(160)}[name]()
n = A.shape[0]
B = ilupp.ILU0Preconditioner(A)
x0 = torch.from_numpy(P.rhs(n, 0)).cuda()
x = x0.clone()
for _ in range(3):
    x.copy_(x0)
    B.apply(x)
torch.cuda.synchronize()
ilupp.set_option("trace", 1)
x.copy_(x0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
B.apply(x)
e1.record()
torch.cuda.synchronize()
print("apply (traced) %.1f us" % (e0.elapsed_time(e1) * 1e3))
np.set_printoptions(linewidth=250, precision=1, suppress=True)
for which in (0, 1):
    dims, tr = B.pr.debug_sys_trace(which)
    ntasks, nper, nty, ntz, twy, twz, _, cfg = dims
    full = np.asarray(tr).reshape(ntz, nty, 3, nper // 3).astype(np.float64)
    nper //= 3
    tr = full[:, :, 0, :].copy()
    wd, wg = full[:, :, 1, -1] / 1965.0, full[:, :, 2, -1] / 1965.0       # us spent waiting for data / for ring inputs
    print("  per tile: wait for TMA data median %.1f us max %.1f ; wait at input gates median %.1f us max %.1f ; tile(0,0): %.1f / %.1f" % (
        np.median(wd), wd.max(), np.median(wg), wg.max(), wd[0, 0], wg[0, 0]))
    t0 = tr[:, :, 0].min()
    tr = (tr - t0) / 1e3          # us
    start, end = tr[:, :, 0], tr[:, :, -1]
    dur = np.diff(tr, axis=2)     # per stage
    print("solve %d: tiles %d x %d (y x z), stages %d, span %.1f us; first tile start 0, last tile start %.1f, last end %.1f" % (
        which, nty, ntz, nper - 1, end.max(), start.max(), end.max()))
    print("  tile duration: median %.1f us min %.1f max %.1f ; stage duration median %.2f us, p90 %.2f, first-stage median %.2f, last-stage median %.2f" % (
        np.median(end - start), (end - start).min(), (end - start).max(), np.median(dur), np.percentile(dur, 90), np.median(dur[:, :, 0]),
        np.median(dur[:, :, -1])))
    print("  stage duration by stage index (median over tiles):", np.median(dur, axis=(0, 1)))
    print("  tile (0,0) stage durations:", dur[0, 0], " tile (1,1):", dur[1, 1])
    print("  start times (us), rows = z tile (every 4th), cols = y tile:")
    print(start[::4, :])
    # lag between neighbouring tiles (start of the third stage), along the first tile row / column (the critical path)
    s2 = tr[:, :, 2]
    print("  y lags along z-tile 0 (us):", np.diff(s2[0, :]))
    print("  z lags along y-tile 0 (us):", np.diff(s2[:, 0]))
    np.save("gpurun_out/tracetile_%s_%d.npy" % (name, which), tr)
