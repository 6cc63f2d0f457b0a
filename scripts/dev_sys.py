"""Bring-up of the line-systolic solve on the GPU box (dev tool): parity of apply / apply_trans against the CPU
oracle for several tilings, then timings.  usage: python scripts/dev_sys.py [parity|time|all]"""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ilupp_b200 as ilupp  # noqa: E402
import oracle  # noqa: E402
from ilupp_b200 import problems as P  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "all"
FAIL = []


def relerr(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def check(name, cond, extra=""):
    print(("  ok   " if cond else "  FAIL ") + name + (" " + str(extra) if extra else ""), flush=True)
    if not cond:
        FAIL.append(name)


def case(name, A, sym, fmts=("csr", "csc")):
    for fmt in fmts:
        M = A.asformat(fmt)
        tag = "%s/%s" % (name, fmt)
        try:
            n = M.shape[0]
            b = P.rhs(n, 0)
            B = ilupp.ILU0Preconditioner(M)
            st = B.stats
            print("%s stats mode=%s ctas=%s grid=%s" % (tag, st["part_mode"], st["nparts"], st["grid"]), flush=True)
            Lo, Uo = oracle.port.ilu0(M)
            for tr in (False, True):
                y = b.copy()
                (B.apply_trans if tr else B.apply)(y)
                yo = oracle.port.apply_lu(Lo, Uo, b, tr)
                check(tag + " ilu0 apply tr=%d" % tr, relerr(y, yo) < 1e-12, "relerr %.2e bitexact=%s" % (relerr(y, yo), np.array_equal(y, yo)))
            if sym:
                C = ilupp.IChol0Preconditioner(M)
                Lco = oracle.port.ichol0(M)
                for tr in (False, True):
                    y = b.copy()
                    (C.apply_trans if tr else C.apply)(y)
                    yo = oracle.port.apply_llt(Lco, b)
                    check(tag + " ichol0 apply tr=%d" % tr, relerr(y, yo) < 1e-12, "relerr %.2e bitexact=%s" % (relerr(y, yo), np.array_equal(y, yo)))
        except Exception:
            traceback.print_exc()
            FAIL.append(tag + " EXCEPTION")


def timing(name, A, kind, reps=20, check_ref=True):
    n = A.shape[0]
    t = time.perf_counter()
    B = (ilupp.IChol0Preconditioner if kind == "ichol0" else ilupp.ILU0Preconditioner)(A)
    dt = time.perf_counter() - t
    st = B.stats
    b = P.rhs(n, 0)
    x = torch.from_numpy(b).cuda()
    x0 = x.clone()
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
    tot = B.total_nnz
    nnzLU = tot if kind == "ilu0" else 2 * tot
    by = 12 * nnzLU + 8 * (n + 1) + 32 * n
    print("%s %s: create %.0f ms (sym %.1f num %.2f pack %.2f) mode %s ctas %s grid %s | apply med %.1f us min %.1f us -> %.0f GB/s (%.1f%% of 6555)" % (
        name, kind, dt * 1e3, *B.timings_ms[:3], st["part_mode"], st["nparts"], st["grid"], np.median(ts) * 1e3, ts.min() * 1e3,
        by / np.median(ts) / 1e6, by / np.median(ts) / 1e6 / 65.55), flush=True)
    if check_ref and oracle.ref is not None:
        R = (oracle.ref.IChol0 if kind == "ichol0" else oracle.ref.ILU0)(A)
        yo = R.apply(b)
        y = x.cpu().numpy()
        e = relerr(y, yo)
        check(name + " " + kind + " vs reference", e < 1e-12, "relerr %.2e bitexact=%s" % (e, np.array_equal(y, yo)))
    return B


print("device:", torch.cuda.get_device_name(0), flush=True)
if mode in ("parity", "all"):
    for exact in (1, 0):
        ilupp.set_option("exact", exact)
        print("=== exact=%d" % exact, flush=True)
        case("p2d7", P.poisson2d(7), True)
        case("p2d40", P.poisson2d(40), True)            # 2 line blocks, one CTA
        case("p2d200", P.poisson2d(200), True, ("csr",))   # 7 line blocks: 2 CTAs chained through the y mailbox
        case("p3d8", P.poisson3d(8), True)              # 8 planes, 1 warp per CTA: every z hand-off remote
        case("cd3d12", P.convdiff3d(12), False)
        case("p3d40", P.poisson3d(40), True, ("csr",))     # 2 line blocks x 40 planes
    ilupp.set_option("exact", 1)
    ilupp.set_option("sys_pz", 3)
    case("p3d20-pz3", P.poisson3d(20), True, ("csr",))
    ilupp.set_option("sys_pz", 0)
    ilupp.set_option("sys_stages", 2)
    case("p3d20-st2", P.poisson3d(20), True, ("csr",))
    ilupp.set_option("sys_stages", 0)
    ilupp.set_option("sys_cluster", 1)
    case("p3d20-nocluster", P.poisson3d(20), True, ("csr",))
    ilupp.set_option("sys_cluster", 3)
    case("p3d20-cluster3", P.poisson3d(20), True, ("csr",))
    ilupp.set_option("sys_cluster", 0)
if mode in ("time", "all"):
    ilupp.set_option("exact", 0)
    timing("p3d128", P.poisson3d(128), "ichol0")
    timing("p3d128", P.poisson3d(128), "ilu0")
    timing("p2d256", P.poisson2d(256), "ilu0")
    timing("p3d64", P.poisson3d(64), "ichol0")
    timing("cd3d160", P.convdiff3d(160), "ilu0")
    ilupp.set_option("exact", 1)
    timing("p3d128-exact", P.poisson3d(128), "ichol0")
print("FAILURES:", FAIL, flush=True)
