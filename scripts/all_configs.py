"""Every BASELINE.json config on the B200 path with the reference's CPU path beside it (bounded samples).

usage: python scripts/all_configs.py [--out gpurun_out/all_configs.json] [--skip-ref] [--small]

Prints one JSON object per config and writes them all to --out.  Definitions as in SURVEY.md 8(d):
apply = both triangular solves, one right-hand side, device-resident, CUDA events, median of 20;
factor = symbolic + numeric + pack from device-resident CSR (first creation pays cudaMalloc; `refactor` is numeric
only on the unchanged pattern); bytes = algorithmic bytes.  Reference = oracle/_ref (the unmodified reference C++),
one host thread; at 160^3 its BiCGStab is timed for a bounded number of iterations and reported per iteration.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ilupp_b200 as ilupp  # noqa: E402
import oracle  # noqa: E402
from ilupp_b200 import problems as P  # noqa: E402

PEAK = 6555.2
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        PEAK = float(json.load(f)["hbm_gbs"])
except Exception:
    pass


def apply_bytes(total_nnz, n, ichol):
    return 12 * (2 * total_nnz if ichol else total_nnz) + 8 * (n + 1) + 32 * n


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def time_apply(B, n, reps=20):
    b = P.rhs(n, 0)
    x0 = torch.from_numpy(b).cuda()
    x = x0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(reps + 3):
        x.copy_(x0)
        torch.cuda.synchronize()
        e0.record()
        B.apply(x)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), b, x.cpu().numpy()


def precond_config(name, A, kind, use_ref, ref_applies=5):
    ichol = kind == "ichol0"
    cls = ilupp.IChol0Preconditioner if ichol else ilupp.ILU0Preconditioner
    n = A.shape[0]
    B = cls(A)                       # first creation (allocator cold)
    del B                            # its blocks go back to the library's cache
    torch.cuda.synchronize()
    t = time.perf_counter()
    B = cls(A)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t) * 1e3
    sym, num, pack, tot = B.timings_ms
    a = torch.from_numpy(A.data).cuda()
    B.refactor(a)
    refac = B.timings_ms[1]
    ms, b, y = time_apply(B, n)
    by = apply_bytes(B.total_nnz, n, ichol)
    st = B.stats
    out = {"config": name, "kind": kind, "n": n, "nnz": int(A.nnz), "total_nnz": int(B.total_nnz),
           "levels": list(st["levels"][:2]), "kernel": {4: "line-systolic", 5: "level loop (deep DAG)", 6: "cooperative level loop"}.get(st["part_mode"], "chunked sync-free"),
           "apply_us": ms * 1e3, "apply_GBs": by / ms / 1e6, "pct_hbm_measured": 100 * by / ms / 1e6 / PEAK,
           "bytes_per_apply": by,
           "factor_ms": {"symbolic": sym, "numeric": num, "pack": pack, "total": tot, "refactor_numeric": refac,
                         "create_wall_incl_h2d": wall}}
    if use_ref and oracle.ref is not None:
        t = time.perf_counter()
        R = oracle.ref.IChol0(A) if ichol else oracle.ref.ILU0(A)
        rf = (time.perf_counter() - t) * 1e3
        yb = b.copy()
        R.apply_inplace(yb)
        ts = []
        for _ in range(ref_applies):
            yy = b.copy()
            t = time.perf_counter()
            R.apply_inplace(yy)
            ts.append(time.perf_counter() - t)
        out["reference_cpu"] = {"cores": 1, "factor_ms": rf, "apply_ms": float(np.median(ts)) * 1e3,
                                "apply_GBs": by / float(np.median(ts)) / 1e9}
        out["apply_relerr_vs_reference"] = relerr(y, yb)
        F, Fr = B.factors(), R.factors()
        out["factor_relerr_vs_reference"] = max(relerr(f.data, g.data) for f, g in zip(F, Fr))
        out["pattern_equal"] = all(np.array_equal(f.indices, g.indices) and np.array_equal(f.indptr, g.indptr)
                                   for f, g in zip(F, Fr))
    print(json.dumps(out), flush=True)
    return out


def bicgstab_config(N, use_ref, ref_iters=3):
    A = P.convdiff3d(N)
    n = A.shape[0]
    b = A @ np.ones(n)
    B = ilupp.ILU0Preconditioner(A)
    x, (it, rel, ab) = B.solve(b, rtol=1e-8, atol=1e-8, max_iter=2000, info=True)   # warm-up (allocations)
    torch.cuda.synchronize()
    t = time.perf_counter()
    xh, (it, rel, ab) = B.solve(b, rtol=1e-8, atol=1e-8, max_iter=2000, info=True)   # host b in, host x out
    dt = time.perf_counter() - t
    nnz_lu = B.total_nnz
    b_apply = apply_bytes(nnz_lu, n, False)
    b_spmv = 12 * A.nnz + 4 * (n + 1) + 16 * n
    b_iter = 2 * (b_apply + b_spmv) + 16 * 8 * n
    out = {"config": "convdiff3d(%d) ILU0 + BiCGStab to 1e-8" % N, "n": n, "nnz": int(A.nnz), "iterations": int(it),
           "rel_residual": float(rel), "abs_residual": float(ab), "solve_ms": dt * 1e3, "ms_per_iteration": dt * 1e3 / max(1, it),
           "bytes_per_iteration": b_iter, "GBs": b_iter * it / dt / 1e9, "pct_hbm_measured": 100 * b_iter * it / dt / 1e9 / PEAK,
           "max_abs_error_vs_ones": float(np.abs(xh - 1).max())}
    if use_ref and oracle.ref is not None:
        R = oracle.ref.ILU0(A)
        t = time.perf_counter()
        try:
            R.bicgstab(A, b, rtol=1e-8, atol=1e-8, max_iter=ref_iters)
        except Exception:
            pass
        dtr = time.perf_counter() - t
        out["reference_cpu"] = {"cores": 1, "ms_per_iteration": dtr * 1e3 / ref_iters,
                                "sample": "%d BiCGStab(SPLIT) iterations of the reference templates with its ILU0" % ref_iters}
    print(json.dumps(out), flush=True)
    return out


def batched_config(nsys, N, nrhs, use_ref):
    """BASELINE config 5: nsys independent 2-D Poisson N x N systems, ILU0 factor + nrhs-column apply"""
    pat, vals = P.batched_poisson2d(nsys, N)
    n1 = pat.shape[0]
    B = ilupp.ILU0PreconditionerBatched(pat, vals)
    del B
    torch.cuda.synchronize()
    B = ilupp.ILU0PreconditionerBatched(pat, vals)
    sym, num, pack, tot = B.timings_ms
    X = torch.randn((nrhs, nsys * n1), dtype=torch.float64, device="cuda").t()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(8):
        torch.cuda.synchronize()
        e0.record()
        B.apply(X)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    by = nsys * (12.0 * B.total_nnz / nsys + 8.0 * (n1 + 1)) + 32.0 * nsys * n1 * nrhs
    out = {"config": "5: %d x poisson2d(%d) ILU0, %d-RHS apply" % (nsys, N, nrhs), "n_per_system": n1, "systems": nsys,
           "kernel": "plane-per-warp (one launch per triangular solve)", "apply_ms": ms, "solves_per_s": nsys * nrhs / ms * 1e3,
           "apply_GBs": by / ms / 1e6, "pct_hbm_measured": 100 * by / ms / 1e6 / PEAK, "bytes_per_apply": by,
           "factor_ms": {"symbolic": sym, "numeric": num, "pack": pack, "total": tot},
           "systems_per_s": nsys / (num + pack + ms) * 1e3}
    if use_ref and oracle.ref is not None:
        As = pat.copy()
        As.data = vals[0].copy()
        t = time.perf_counter()
        R = oracle.ref.ILU0(As)
        rf = time.perf_counter() - t
        b = P.rhs(n1, 1)
        ts = []
        for _ in range(nrhs):
            yy = b.copy()
            t = time.perf_counter()
            R.apply_inplace(yy)
            ts.append(time.perf_counter() - t)
        ra = float(np.median(ts))
        out["reference_cpu"] = {"cores": 1, "sample": "one system: factorisation + %d applies, scaled to %d systems" % (nrhs, nsys),
                                "factor_ms_per_system": rf * 1e3, "apply_ms_per_solve": ra * 1e3,
                                "solves_per_s": 1.0 / ra, "systems_per_s": 1.0 / (rf + nrhs * ra)}
    print(json.dumps(out), flush=True)
    return out


def factors_config(N, use_ref):
    """the reference's threshold factorisations stay on the CPU (compiled reference, oracle/_ref);
    their factors are applied on the GPU through FactorsPreconditioner"""
    m = oracle.ref_module()
    if m is None:
        return []
    A = P.poisson3d(N)
    A.indices = A.indices.astype(np.int32)
    A.indptr = A.indptr.astype(np.int32)
    n = A.shape[0]
    res = []
    for name, make in (("ILUT(fill 10, 1e-3)", lambda: m.ILUTPreconditioner(A.data, A.indices, A.indptr, True, 10, 1e-3)),
                       ("ICholT(fill 5, 1e-3)", lambda: m.ICholTPreconditioner(A.data, A.indices, A.indptr, True, 5, 1e-3))):
        t = time.perf_counter()
        R = make()
        rf = (time.perf_counter() - t) * 1e3
        fs = [(sp.csr_matrix if f[3] else sp.csc_matrix)((f[0], f[1], f[2]), shape=(n, n)) for f in R.factors_info()]
        t = time.perf_counter()
        B = ilupp.FactorsPreconditioner(fs[0], fs[1] if len(fs) > 1 else None)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t) * 1e3
        ms, b, y = time_apply(B, n)
        ts = []
        for _ in range(3):
            yy = b.copy()
            t = time.perf_counter()
            R.apply(yy)
            ts.append(time.perf_counter() - t)
        st = B.stats
        by = apply_bytes(B.total_nnz, n, len(fs) == 1)
        out = {"config": "f1: poisson3d(%d) %s factors from the reference, GPU apply" % (N, name), "n": n,
               "total_nnz": int(B.total_nnz), "levels": list(st["levels"][:2]),
               "kernel": {4: "line-systolic", 5: "level loop (deep DAG)", 6: "cooperative level loop"}.get(st["part_mode"], "chunked sync-free"),
               "apply_us": ms * 1e3, "apply_GBs": by / ms / 1e6, "pct_hbm_measured": 100 * by / ms / 1e6 / PEAK,
               "create_wall_ms_incl_host_assembly": wall, "symbolic_ms": B.timings_ms[0],
               "reference_cpu": {"cores": 1, "factor_ms": rf, "apply_ms": float(np.median(ts)) * 1e3},
               "apply_relerr_vs_reference": relerr(y, yy)}
        print(json.dumps(out), flush=True)
        res.append(out)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/all_configs.json")
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--small", action="store_true", help="reduced sizes (smoke run of this script)")
    args = ap.parse_args()
    use_ref = not args.skip_ref
    ilupp.set_option("exact", 0)
    res = []
    print("device:", torch.cuda.get_device_name(0), "| host cores:", os.cpu_count(), "| reference built:", oracle.ref is not None, flush=True)
    res.append(precond_config("0: poisson2d(256) ILU0 csr", P.poisson2d(64 if args.small else 256), "ilu0", use_ref))
    res.append(precond_config("0: poisson2d(256) ILU0 csc", P.poisson2d(64 if args.small else 256).tocsc(), "ilu0", use_ref))
    res.append(precond_config("1: poisson3d(128) IChol0", P.poisson3d(24 if args.small else 128), "ichol0", use_ref))
    res.append(precond_config("1b: poisson3d(128) ILU0", P.poisson3d(24 if args.small else 128), "ilu0", use_ref))
    res.append(precond_config("3a: convdiff3d(160) ILU0", P.convdiff3d(20 if args.small else 160), "ilu0", use_ref, ref_applies=3))
    res.append(bicgstab_config(20 if args.small else 160, use_ref))
    A = P.irregular_powerlaw(20000 if args.small else 1_000_000)
    rl = np.diff(A.indptr)
    o = precond_config("4: irregular power-law n=1M", A, "ilu0", use_ref, ref_applies=3)
    o["row_length"] = {"max": int(rl.max()), "mean": float(rl.mean())}
    res.append(o)
    res.append(batched_config(64 if args.small else 4096, 128, 16, use_ref))
    if use_ref:
        res += factors_config(16 if args.small else 96, use_ref)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
