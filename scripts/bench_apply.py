"""Device-resident apply timing for one problem under a set of libilub options (dev tool).

usage: python scripts/bench_apply.py <problem> <kind> [key=value ...] [reps=N] [check=1]
problems: p2d256 p3d64 p3d128 cd3d160 irr1m   kinds: ilu0 ichol0
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ilupp_b200 as ilupp  # noqa: E402
from ilupp_b200 import problems as P  # noqa: E402

PROBLEMS = {
    "p2d128": lambda: P.poisson2d(128),
    "p2d256": lambda: P.poisson2d(256),
    "p3d32": lambda: P.poisson3d(32),
    "p3d64": lambda: P.poisson3d(64),
    "p3d128": lambda: P.poisson3d(128),
    "cd3d160": lambda: P.convdiff3d(160),
    "irr1m": lambda: P.irregular_powerlaw(),
    "irr100k": lambda: P.irregular_powerlaw(100000),
}


def apply_bytes(B, kind, n):
    tot = B.total_nnz
    nnzLU = tot if kind == "ilu0" else 2 * tot
    return 12 * nnzLU + 8 * (n + 1) + 32 * n


def main():
    prob, kind = sys.argv[1], sys.argv[2]
    reps, check = 20, 0
    for kv in sys.argv[3:]:
        k, v = kv.split("=")
        if k == "reps":
            reps = int(v)
        elif k == "check":
            check = int(v)
        else:
            ilupp.set_option(k, int(v))
    A = PROBLEMS[prob]()
    n = A.shape[0]
    t = time.perf_counter()
    B = (ilupp.IChol0Preconditioner if kind == "ichol0" else ilupp.ILU0Preconditioner)(A)
    dt = time.perf_counter() - t
    b = P.rhs(n, 0)
    x0 = torch.from_numpy(b).cuda()
    x = x0.clone()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(reps + 3):
        x.copy_(x0)
        torch.cuda.synchronize()
        ev0.record()
        B.apply(x)
        ev1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(ev0.elapsed_time(ev1))
    ts = np.array(ts)
    by = apply_bytes(B, kind, n)
    med = float(np.median(ts))
    print("%s %s %s | create %.0f ms (sym %.1f num %.2f pack %.2f) | apply med %.1f us min %.1f us | %.0f GB/s = %.1f%% of 6555 | levels %s parts %s mode %s grid %s" % (
        prob, kind, " ".join(sys.argv[3:]), dt * 1e3, *B.timings_ms[:3], med * 1e3, ts.min() * 1e3, by / med / 1e6,
        by / med / 1e6 / 65.55, B.stats["levels"][:2], B.stats["nparts"], B.stats["part_mode"], B.stats["grid"]), flush=True)
    if check:
        import oracle
        R = (oracle.ref.IChol0 if kind == "ichol0" else oracle.ref.ILU0)(A)
        yo = R.apply(b)
        y = x.cpu().numpy()
        print("   parity vs reference: relerr %.2e bitexact=%s" % (np.abs(y - yo).max() / np.abs(yo).max(), np.array_equal(y, yo)), flush=True)


if __name__ == "__main__":
    main()
