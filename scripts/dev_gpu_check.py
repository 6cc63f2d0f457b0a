"""Bring-up script for the GPU box (not a test): parity of every stage against the CPU oracle on small
inputs, then first timings.  Usage: python scripts/dev_gpu_check.py [quick|full]"""
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

mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
FAIL = []


def relerr(a, b):
    d = np.max(np.abs(a - b))
    s = max(np.max(np.abs(b)), 1e-300)
    return d / s


def same_pattern(X, Y):
    return X.format == Y.format and np.array_equal(X.indptr, Y.indptr) and np.array_equal(X.indices, Y.indices)


def check(name, cond, extra=""):
    print(("  ok   " if cond else "  FAIL ") + name + (" " + str(extra) if extra else ""), flush=True)
    if not cond:
        FAIL.append(name)


def case(name, A, sym):
    for fmt in ("csr", "csc"):
        M = A.asformat(fmt)
        tag = "%s/%s" % (name, fmt)
        try:
            n = M.shape[0]
            b = P.rhs(n, 0)
            t = time.perf_counter()
            B = ilupp.ILU0Preconditioner(M)
            dt = time.perf_counter() - t
            print("%s ILU0 create %.1f ms, timings %s stats %s" % (tag, dt * 1e3, B.timings_ms, B.stats), flush=True)
            L, U = B.factors()
            Lo, Uo = oracle.port.ilu0(M)
            check(tag + " ilu0 pattern", same_pattern(L, Lo) and same_pattern(U, Uo))
            check(tag + " ilu0 values bitexact", np.array_equal(L.data, Lo.data) and np.array_equal(U.data, Uo.data),
                  "relerr L %.2e U %.2e" % (relerr(L.data, Lo.data), relerr(U.data, Uo.data)))
            check(tag + " total_nnz", B.total_nnz == Lo.nnz + Uo.nnz)
            for tr in (False, True):
                y = b.copy()
                (B.apply_trans if tr else B.apply)(y)
                yo = oracle.port.apply_lu(Lo, Uo, b, tr)
                check(tag + " ilu0 apply tr=%d" % tr, relerr(y, yo) < 1e-12, "relerr %.2e bitexact=%s" % (relerr(y, yo), np.array_equal(y, yo)))
            # levels
            Mc = M if fmt == "csr" else M.T.tocsr()  # CSR-read arrays
            for which in range(4):
                lev, nlev = B.levels(which)
                gather = (fmt == "csr") == (which < 2)
                lower = (which % 2 == 0)
                if gather:
                    lo, nlo = oracle.port.levels(Mc.indptr, Mc.indices, lower=lower)
                else:
                    # scatter-form solves: first = U^T forward (entries of strict upper), second = L^T backward
                    lo, nlo = oracle.port.levels(Mc.indptr, Mc.indices, lower=lower, transposed=True)
                check(tag + " levels which=%d" % which, nlev == nlo and np.array_equal(lev, lo), "nlev %d/%d" % (nlev, nlo))
            # spmv
            x = P.rhs(n, 3)
            check(tag + " spmv", relerr(B.pr.spmv(x), oracle.port.spmv(M, x)) < 1e-14, "relerr %.2e" % relerr(B.pr.spmv(x), oracle.port.spmv(M, x)))
            # multi-RHS
            X = P.rhs(n, 5, k=3)
            Y = B @ X
            ok = all(relerr(Y[:, c], oracle.port.apply_lu(Lo, Uo, X[:, c])) < 1e-12 for c in range(3))
            check(tag + " multi-rhs", ok)
            # bicgstab
            bb = M @ np.ones(n)
            xs, (it, r, a) = B.solve(bb, 1e-8, 1e-8, 2000, info=True)
            xo = oracle.port.bicgstab(M, bb, (Lo, Uo))
            check(tag + " bicgstab", abs(it - xo[1]) <= 2 and np.abs(xs - 1).max() < 1e-6, "iters %d/%d err %.2e" % (it, xo[1], np.abs(xs - 1).max()))
            if sym:
                C = ilupp.IChol0Preconditioner(M)
                Lc, = C.factors()
                Lco = oracle.port.ichol0(M)
                check(tag + " ichol0 pattern", same_pattern(Lc, Lco))
                check(tag + " ichol0 values bitexact", np.array_equal(Lc.data, Lco.data), "relerr %.2e" % relerr(Lc.data, Lco.data))
                for tr in (False, True):
                    y = b.copy()
                    (C.apply_trans if tr else C.apply)(y)
                    yo = oracle.port.apply_llt(Lco, b)
                    check(tag + " ichol0 apply tr=%d" % tr, relerr(y, yo) < 1e-12, "relerr %.2e bitexact=%s" % (relerr(y, yo), np.array_equal(y, yo)))
        except Exception:
            traceback.print_exc()
            FAIL.append(tag + " EXCEPTION")


def timing(name, A, kind, reps=20):
    n = A.shape[0]
    t = time.perf_counter()
    B = (ilupp.IChol0Preconditioner if kind == "ichol0" else ilupp.ILU0Preconditioner)(A)
    dt = time.perf_counter() - t
    st = B.stats
    print("%s %s: n=%d nnz=%d create %.1f ms (symbolic %.2f numeric %.2f pack %.2f) levels %s" % (
        name, kind, n, A.nnz, dt * 1e3, *B.timings_ms[:3], st["levels"]), flush=True)
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
    bytes_ = 12 * nnzLU + 8 * (n + 1) + 32 * n
    print("   apply: median %.1f us  min %.1f us  -> %.1f GB/s algorithmic (%.1f%% of 6555)" % (
        np.median(ts) * 1e3, ts.min() * 1e3, bytes_ / np.median(ts) / 1e6, bytes_ / np.median(ts) / 1e6 / 65.55), flush=True)
    y = x.cpu().numpy()
    if kind == "ichol0":
        R = oracle.ref.IChol0(A) if oracle.ref else None
    else:
        R = oracle.ref.ILU0(A) if oracle.ref else None
    if R is not None:
        t = time.perf_counter()
        yo = R.apply(b)
        dtc = time.perf_counter() - t
        print("   vs reference: relerr %.2e bitexact=%s (ref apply %.1f ms)" % (relerr(y, yo), np.array_equal(y, yo), dtc * 1e3), flush=True)
        F = B.factors()
        Fr = R.factors()
        for a_, b_ in zip(F, Fr):
            print("   factor: pattern %s values relerr %.2e bitexact=%s" % (same_pattern(a_, b_), relerr(a_.data, b_.data), np.array_equal(a_.data, b_.data)), flush=True)
    return B


print("device:", torch.cuda.get_device_name(0), flush=True)
case("tridiag50", P.poisson1d(50), True)
case("p2d16", P.poisson2d(16), True)
case("p3d8", P.poisson3d(8), True)
case("cd3d8", P.convdiff3d(8), False)
case("irr3000", P.irregular_powerlaw(3000, 6, seed=1), False)
case("irr20000w", P.irregular_powerlaw(20000, 19, seed=2), False)
print("---- level-scheduled variants", flush=True)
ilupp.set_option("solve_algo", 1)
ilupp.set_option("factor_algo", 1)
case("LS p3d8", P.poisson3d(8), True)
case("LS irr3000", P.irregular_powerlaw(3000, 6, seed=1), False)
ilupp.set_option("solve_algo", 0)
ilupp.set_option("factor_algo", 0)
print("---- inexact arithmetic", flush=True)
ilupp.set_option("exact", 0)
case("FMA p3d8", P.poisson3d(8), True)
ilupp.set_option("exact", 1)
print("FAILURES:", FAIL, flush=True)
if mode != "quick" or not FAIL:
    timing("p2d256", P.poisson2d(256), "ilu0")
    timing("p3d64", P.poisson3d(64), "ichol0")
    timing("p3d128", P.poisson3d(128), "ichol0")
    timing("p3d128", P.poisson3d(128), "ilu0")
    ilupp.set_option("exact", 0)
    timing("p3d128 fma", P.poisson3d(128), "ichol0")
    ilupp.set_option("exact", 1)
print("FAILURES:", FAIL, flush=True)
