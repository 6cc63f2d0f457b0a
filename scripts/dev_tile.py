"""Bring-up of the tile-systolic solve (ilub_tile.cu) on the GPU box (dev tool): parity against the CPU oracle on
ragged / small / full-size grids in both arithmetic modes, then timings.
usage: python scripts/dev_tile.py [parity|time|all] [key=value options ...]"""
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
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ilupp.set_option(k, int(v))
FAIL = []


def relerr(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def check(name, cond, extra=""):
    print(("  ok   " if cond else "  FAIL ") + name + (" " + str(extra) if extra else ""), flush=True)
    if not cond:
        FAIL.append(name)


def grid3d(nx, ny, nz, seed=0, sym=False):
    """7-point stencil on an nx x ny x nz grid (x fastest), random coefficients, diagonally dominant"""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    n = nx * ny * nz
    i = np.arange(n)
    ix, iy, iz = i % nx, (i // nx) % ny, i // (nx * ny)
    rows, cols = [], []
    for off, ok in ((-1, ix > 0), (1, ix < nx - 1), (-nx, iy > 0), (nx, iy < ny - 1), (-nx * ny, iz > 0), (nx * ny, iz < nz - 1)):
        rows.append(i[ok])
        cols.append(i[ok] + off)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    A = sp.csr_matrix((-rng.uniform(0.5, 1.5, rows.size), (rows, cols)), shape=(n, n))
    if sym:
        A = (A + A.T) * 0.5
    A = (A + sp.diags(1.0 + np.asarray(abs(A).sum(axis=1)).ravel())).tocsr()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32)
    A.indptr = A.indptr.astype(np.int32)
    return A


def case(name, A, sym, fmts=("csr", "csc")):
    for exact in (1, 0):
        ilupp.set_option("exact", exact)
        for fmt in fmts:
            M = A.asformat(fmt)
            tag = "%s/%s/exact%d" % (name, fmt, exact)
            try:
                n = M.shape[0]
                b = P.rhs(n, 0)
                B = ilupp.ILU0Preconditioner(M)
                st = B.stats
                print("%s stats mode=%s ctas=%s grid=%s" % (tag, st["part_mode"], st["nparts"], st["grid"]), flush=True)
                Lo, Uo = oracle.port.ilu0(M)
                L, U = B.factors()
                check(tag + " factors", relerr(L.data, Lo.data) < 1e-12 and relerr(U.data, Uo.data) < 1e-12)
                for tr in (False, True):
                    y = b.copy()
                    (B.apply_trans if tr else B.apply)(y)
                    yo = oracle.port.apply_lu(Lo, Uo, b, tr)
                    ok = np.array_equal(y, yo) if exact else relerr(y, yo) < 1e-12
                    check(tag + " ilu0 apply tr=%d" % tr, ok, "relerr %.2e bitexact=%s" % (relerr(y, yo), np.array_equal(y, yo)))
                if sym:
                    C = ilupp.IChol0Preconditioner(M)
                    Lco = oracle.port.ichol0(M)
                    (Lc,) = C.factors()
                    check(tag + " ichol factor", relerr(Lc.data, Lco.data) < 1e-12)
                    for tr in (False, True):
                        y = b.copy()
                        (C.apply_trans if tr else C.apply)(y)
                        yo = oracle.port.apply_llt(Lco, b)
                        ok = np.array_equal(y, yo) if exact else relerr(y, yo) < 1e-12
                        check(tag + " ichol0 apply tr=%d" % tr, ok, "relerr %.2e bitexact=%s" % (relerr(y, yo), np.array_equal(y, yo)))
            except Exception:
                traceback.print_exc()
                FAIL.append(tag + " EXCEPTION")


def timing(name, A, kind, reps=20):
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
    print("%s %s: create %.0f ms (sym %.1f num %.2f pack %.2f) mode %s ctas %s grid %s | apply med %.1f us min %.1f us -> %.0f GB/s (%.1f%% of 6458)" % (
        name, kind, dt * 1e3, *B.timings_ms[:3], st["part_mode"], st["nparts"], st["grid"], np.median(ts) * 1e3, ts.min() * 1e3,
        by / np.median(ts) / 1e6, by / np.median(ts) / 1e6 / 64.58), flush=True)
    if oracle.ref is not None:
        R = (oracle.ref.IChol0 if kind == "ichol0" else oracle.ref.ILU0)(A)
        yo = R.apply(b)
        y = x.cpu().numpy()
        check(name + " " + kind + " vs reference", relerr(y, yo) < 1e-12, "relerr %.2e" % relerr(y, yo))
    return B


if mode in ("parity", "all"):
    case("g3d 16x8x4 (one tile)", grid3d(16, 8, 4, 1), False)
    case("g3d 10x9x5 ragged", grid3d(10, 9, 5, 2), False)
    case("g3d 24^3 sym", grid3d(24, 24, 24, 3, sym=True), True, fmts=("csr",))
    case("g3d 30x21x13 ragged", grid3d(30, 21, 13, 4), False)
    case("g3d 6x40x37", grid3d(6, 40, 37, 5), False, fmts=("csr",))
    case("g3d 64x50x30", grid3d(64, 50, 30, 6), False, fmts=("csr",))
    case("poisson3d 48", P.poisson3d(48), True, fmts=("csr",))
    print("parity failures:", FAIL, flush=True)

if mode in ("time", "all"):
    for ex in (0, 1):
        ilupp.set_option("exact", ex)
        print("exact =", ex)
        timing("poisson3d 128", P.poisson3d(128), "ichol0")
        timing("poisson3d 128", P.poisson3d(128), "ilu0")
        timing("convdiff3d 160", P.convdiff3d(160), "ilu0")
        timing("poisson3d 64", P.poisson3d(64), "ichol0")
    print("failures:", FAIL, flush=True)
sys.exit(1 if FAIL else 0)
