"""dev: block-level timeline of the line-systolic solve (option trace).  usage: trace_sys.py <problem> [key=value ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("ILUB_PKG_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P

name = sys.argv[1]
for kv in sys.argv[2:]:
    k, v = kv.split("="); ilupp.set_option(k, int(v))
ilupp.set_option("exact", 0)
A = {"p2d256": lambda: P.poisson2d(256), "p3d128": lambda: P.poisson3d(128), "p3d64": lambda: P.poisson3d(64),
     "cd3d160": lambda: P.convdiff3d(160)}[name]()
n = A.shape[0]
B = ilupp.ILU0Preconditioner(A)
x0 = torch.from_numpy(P.rhs(n, 0)).cuda(); x = x0.clone()
for _ in range(3):
    x.copy_(x0); B.apply(x)
torch.cuda.synchronize()
ilupp.set_option("trace", 1)
x.copy_(x0); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); B.apply(x); e1.record(); torch.cuda.synchronize()
print("apply (traced) %.1f us" % (e0.elapsed_time(e1) * 1e3))
for which in (0, 1):
    dims, tr = B.pr.debug_sys_trace(which)
    ntasks, nper, nyb, nplanes, pyb, pz, csize, cfg = dims
    full = np.asarray(tr).reshape(ntasks, 3, nper // 3).astype(np.float64)
    np.save("gpurun_out/tracefull_%s_%d.npy" % (name, which), full)
    nper //= 3
    tr = full[:, 0, :].copy()
    sp = full[:, 1, :]
    wc = full[:, 2, :]
    t0 = tr[:, 0].min()
    tr -= t0
    print("  failed polls per task: median %.0f max %.0f ; per plane-group position (wz): %s" % (
        np.median(sp[:, -1]), sp[:, -1].max(), [float(np.median(sp[:, -1].reshape(nplanes, nyb)[wz::pz, :])) for wz in range(pz)]))
    print("  helper/stage wait per task (us): median %.1f max %.1f" % (np.median(wc[:, -1]) / 1965, wc[:, -1].max() / 1965))
    start, end = tr[:, 0], tr[:, -1]
    print("solve %d: tasks %d (nyb %d x planes %d), tile %dx%d, cluster %d, stages/rz %d | span %.1f us" % (
        which, ntasks, nyb, nplanes, pyb, pz, csize, cfg, end.max() / 1e3))
    dur = np.diff(tr, axis=1)
    print("  per-block duration (ns): median %.0f p10 %.0f p90 %.0f max %.0f -> %.0f ns/step" % (
        np.median(dur), np.percentile(dur, 10), np.percentile(dur, 90), dur.max(), np.median(dur) / 8))
    T3 = tr.reshape(nplanes, nyb, nper)
    print("  plane end-time slope (jb 0): %.0f ns/plane ; task duration plane 0: %.1f us, last plane: %.1f us" % (
        (T3[-1, 0, -1] - T3[0, 0, -1]) / max(1, nplanes - 1), (T3[0, 0, -1] - T3[0, 0, 0]) / 1e3, (T3[-1, 0, -1] - T3[-1, 0, 0]) / 1e3))
    S = start.reshape(nplanes, nyb)
    E = end.reshape(nplanes, nyb)
    print("  task start (us) by plane for line block 0:", np.round(S[::max(1, nplanes // 16), 0] / 1e3, 1))
    print("  task start (us) by line block for plane 0:", np.round(S[0, :] / 1e3, 1))
    print("  task end   (us) by plane for last line block:", np.round(E[::max(1, nplanes // 16), -1] / 1e3, 1))
    # z hand-off lag between consecutive planes at mid-task (block nper//2)
    np.save("gpurun_out/trace_%s_%d.npy" % (name, which), tr)
    mid = tr[:, nper // 2].reshape(nplanes, nyb)[:, 0]
    d = np.diff(mid)
    if d.size == 0:
        continue
    print("  plane-to-plane lag at mid task (ns): median %.0f max %.0f ; every pz-th: %s" % (np.median(d), d.max(), np.round(d[pz - 1::pz][:12])))
