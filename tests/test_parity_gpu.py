"""GPU parity tests (pytest -m gpu): the CUDA path through the public API / C ABI against
  - the committed golden fixtures from the unmodified reference (tests/golden),
  - the CPU oracle (oracle/ilub_oracle.c) on seeded inputs,
  - the reference's own hot-path test cases (test/tests.py),
  - size-independent properties at the BASELINE.json sizes.
Bars (BASELINE.json north_star): sparsity pattern and level sets bit-exact; factor values and apply()
outputs within 1e-12 relative (max norm).  With option exact=1 (default) lane-per-row paths are
additionally bit-for-bit equal to the reference compiled without FMA contraction."""
import numpy as np
import pytest
import scipy.sparse

import oracle
from ilupp_b200 import problems as P
from helpers import (example_laplace, golden_factor, golden_matrix, ichol0_dense, ilu0_dense, laplace2d_matrix,
                     laplace_matrix, load_golden, random_matrix, relerr, same_pattern)

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ilupp():
    import ilupp_b200
    assert ilupp_b200._ilupp.device_count() > 0, "no CUDA device: the gpu tests must run on the B200 box"
    ilupp_b200.set_option("exact", 1)
    ilupp_b200.set_option("solve_algo", 0)
    ilupp_b200.set_option("factor_algo", 0)
    return ilupp_b200


CASES = [(p, f) for p in ("laplace", "laplace2d", "random", "nonsym_tridiag") for f in ("csr", "csc")]


# ---- golden fixtures ---------------------------------------------------------------------------------
@pytest.mark.parametrize("prob,fmt", CASES)
def test_golden_ilu0(ilupp, prob, fmt):
    g = load_golden("%s_%s" % (prob, fmt))
    A = golden_matrix(g)
    B = ilupp.ILU0Preconditioner(A)
    L, U = B.factors()
    Lg, Ug = golden_factor(g, "ilu0_L"), golden_factor(g, "ilu0_U")
    assert same_pattern(L, Lg) and same_pattern(U, Ug)
    assert relerr(L.data, Lg.data) <= TOL and relerr(U.data, Ug.data) <= TOL
    assert np.array_equal(L.data, Lg.data) and np.array_equal(U.data, Ug.data)
    assert B.total_nnz == int(g["ilu0_total_nnz"])
    y = g["b"].copy()
    B.apply(y)
    assert relerr(y, g["ilu0_apply"]) <= TOL and np.array_equal(y, g["ilu0_apply"])
    y = g["b"].copy()
    B.apply_trans(y)
    assert relerr(y, g["ilu0_apply_trans"]) <= TOL and np.array_equal(y, g["ilu0_apply_trans"])


@pytest.mark.parametrize("prob,fmt", CASES)
def test_golden_ichol0(ilupp, prob, fmt):
    g = load_golden("%s_%s" % (prob, fmt))
    S = golden_matrix(g, "S")
    B = ilupp.IChol0Preconditioner(S)
    (L,) = B.factors()
    Lg = golden_factor(g, "ichol0_L")
    assert L.format == "csr" and same_pattern(L, Lg)
    assert np.array_equal(L.data, Lg.data, equal_nan=True)
    assert B.total_nnz == int(g["ichol0_total_nnz"])
    for fn in (B.apply, B.apply_trans):
        y = g["b"].copy()
        fn(y)
        assert np.array_equal(y, g["ichol0_apply"], equal_nan=True)


def test_golden_bicgstab(ilupp):
    g = load_golden("bicgstab_convdiff12")
    n = g["A_indptr"].size - 1
    A = scipy.sparse.csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=(n, n))
    x, (it, rel, ab) = ilupp.solve(A, g["b"], rtol=1e-8, atol=1e-8, max_iter=2000, info=True)
    # parallel reductions reorder the dot products: iteration count within +-2, solution to solver tolerance
    assert abs(it - int(g["iters"])) <= 2
    assert relerr(x, g["x"]) < 1e-7 and rel < 1e-8 and ab < 1e-8


# ---- the reference's own hot-path tests (test/tests.py) through the drop-in API -------------------------
@pytest.mark.parametrize("cls", ["ILU0Preconditioner", "IChol0Preconditioner"])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_ref_solve_in_one_step(ilupp, cls, fmt):  # tests.py:197-207,286-299
    A, b, x_exact = example_laplace(50, fmt)
    Pr = getattr(ilupp, cls)(A)
    x = b.copy()
    Pr.apply(x)
    assert np.allclose(x, x_exact)
    assert np.allclose(Pr.T @ (A.T @ x_exact), x_exact)
    assert np.allclose(Pr @ b, x_exact) and np.allclose(Pr.dot(b), x_exact)


@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_ref_factors_correct(ilupp, fmt):  # tests.py:229-259
    A = laplace_matrix(50, fmt)
    L, U = ilupp.ILU0Preconditioner(A).factors()
    assert all(i >= j for i, j in zip(*L.nonzero())) and all(i <= j for i, j in zip(*U.nonzero()))
    assert np.allclose(A.toarray(), (L @ U).toarray())
    (Lc,) = ilupp.IChol0Preconditioner(A).factors()
    assert np.allclose(A.toarray(), (Lc @ Lc.T).toarray())


def test_ref_total_nnz_and_repr(ilupp):  # tests.py:301-306 + SURVEY appendix A
    A = laplace_matrix(50)
    B = ilupp.ILU0Preconditioner(A)
    assert B.total_nnz == 2 * (2 * 50 - 1)
    assert ilupp.IChol0Preconditioner(A).total_nnz == 2 * 50 - 1
    assert repr(B) == "<50x50 ILU0Preconditioner with nnz=198, dtype=float64>"
    assert B.pr.exists is True and B.pr.special_info == "" and B.pr.memory == 0.0
    assert B.shape == (50, 50)


@pytest.mark.parametrize("gen", [laplace_matrix, laplace2d_matrix, random_matrix])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_ref_standalone_factorisations(ilupp, gen, fmt):  # tests.py:324-338
    A = gen(50, fmt)
    X, Xr = ilupp.ilu0(A), ilu0_dense(A)
    assert all(Xi.format == fmt for Xi in X)
    assert all(np.allclose(Xi.toarray(), Ri) for Xi, Ri in zip(X, Xr))
    S = ((A + A.T) / 2).asformat(fmt)
    Lc = ilupp.ichol0(S)
    assert Lc.format == "csr" and np.allclose(Lc.toarray(), ichol0_dense(S))


def test_layout_conventions(ilupp):  # explicit unit diagonals and their position
    A = laplace_matrix(50, "csr")
    L, U = ilupp.ilu0(A)
    assert np.array_equal(L.indices[L.indptr[1]:L.indptr[2]], [0, 1])      # unit diagonal LAST
    assert np.array_equal(U.indices[U.indptr[1]:U.indptr[2]], [1, 2])      # diagonal FIRST
    assert L.data[L.indptr[2] - 1] == 1.0
    Lc, Uc = ilupp.ilu0(A.tocsc())
    assert Lc.format == "csc" and Uc.format == "csc"
    assert not np.allclose(Lc.diagonal(), 1.0) and np.allclose(Uc.diagonal(), 1.0)   # ILU0.hpp:100-105
    assert L.has_sorted_indices and U.has_sorted_indices


# ---- seeded inputs against the CPU oracle ---------------------------------------------------------------
PROBLEMS = [
    ("poisson2d_40", lambda: P.poisson2d(40), True),
    ("poisson3d_13", lambda: P.poisson3d(13), True),
    ("convdiff3d_12", lambda: P.convdiff3d(12), False),
    ("powerlaw_6000", lambda: P.irregular_powerlaw(6000, 7, seed=3), False),
    ("powerlaw_wide_12000", lambda: P.irregular_powerlaw(12000, 19, seed=4), False),   # warp-cooperative chunks
    ("dense_small_24", lambda: scipy.sparse.csr_matrix(np.random.default_rng(0).random((24, 24)) + 24 * np.eye(24)), False),
    ("one_by_one", lambda: scipy.sparse.csr_matrix(np.array([[2.5]])), True),
    ("diagonal_70", lambda: scipy.sparse.diags([np.arange(1.0, 71.0)], [0]).tocsr(), True),
]


@pytest.mark.parametrize("name,gen,sym", PROBLEMS, ids=[p[0] for p in PROBLEMS])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_oracle_parity(ilupp, name, gen, sym, fmt):
    A = gen().asformat(fmt)
    A.indices = A.indices.astype(np.int32)
    A.indptr = A.indptr.astype(np.int32)
    n = A.shape[0]
    b = P.rhs(n, 2)
    B = ilupp.ILU0Preconditioner(A)
    L, U = B.factors()
    Lo, Uo = oracle.port.ilu0(A)
    assert same_pattern(L, Lo) and same_pattern(U, Uo)
    assert relerr(L.data, Lo.data) <= TOL and relerr(U.data, Uo.data) <= TOL
    assert np.array_equal(L.data, Lo.data) and np.array_equal(U.data, Uo.data)
    assert B.total_nnz == Lo.nnz + Uo.nnz
    for tr in (False, True):
        y = b.copy()
        (B.apply_trans if tr else B.apply)(y)
        assert relerr(y, oracle.port.apply_lu(Lo, Uo, b, tr)) <= TOL
    # level sets: bit-exact against the CPU restatement (no reference counterpart)
    Mc = A if fmt == "csr" else A.T.tocsr()
    Mc.sort_indices()
    for which in range(4):
        lev, nlev = B.levels(which)
        gather = (fmt == "csr") == (which < 2)
        lo, nlo = oracle.port.levels(Mc.indptr, Mc.indices, lower=(which % 2 == 0), transposed=not gather)
        assert nlev == nlo and np.array_equal(lev, lo)
        lptr, lrows = B.level_sets(which)
        order = np.lexsort((np.arange(n), lo))                      # canonical listing: (level, row)
        assert np.array_equal(lrows, order) and lptr[0] == 0 and lptr[-1] == n
        assert np.array_equal(np.diff(lptr), np.bincount(lo, minlength=nlo))
    assert relerr(B.pr.spmv(b), oracle.port.spmv(A, b)) <= 1e-14
    X = P.rhs(n, 5, k=4)
    Y = B @ X                                                       # native multi-RHS
    for c in range(4):
        assert relerr(Y[:, c], oracle.port.apply_lu(Lo, Uo, X[:, c])) <= TOL
    if sym:
        Cc = ilupp.IChol0Preconditioner(A)
        (Lc,) = Cc.factors()
        Lco = oracle.port.ichol0(A)
        assert same_pattern(Lc, Lco) and np.array_equal(Lc.data, Lco.data)
        y = b.copy()
        Cc.apply(y)
        assert relerr(y, oracle.port.apply_llt(Lco, b)) <= TOL


@pytest.mark.parametrize("opts", [{"solve_algo": 1, "factor_algo": 1}, {"exact": 0}, {"solve_algo": 2},
                                  {"solve_algo": 2, "exact": 0}, {"solve_algo": 2, "warps_per_cta": 4, "nparts": 3},
                                  {"solve_algo": 2, "partition": 1}, {"solve_algo": 2, "stages": 1, "wcap": 2},
                                  {"solve_algo": 2, "ring_chunks": 64, "near_window": 8}, {"sys_pz": 3},
                                  {"sys_stages": 2, "sys_pz": 2}, {"factor_algo": 2}, {"factor_algo": 3},
                                  {"factor_algo": 3, "exact": 0}, {"solve_algo": 3}, {"solve_algo": 3, "exact": 0},
                                  {"solve_algo": 4}, {"solve_algo": 4, "exact": 0},
                                  {"solve_algo": 4, "levels_cluster": 0}, {"solve_algo": 4, "levels_cluster": 0, "exact": 0}],
                         ids=["level-scheduled", "fma", "syncfree", "syncfree-fma", "few-warps", "contiguous",
                              "one-stage-narrow", "small-ring", "systolic-pz3", "systolic-2stage", "factor-syncfree",
                              "factor-level-loop", "factor-level-loop-fma", "solve-level-loop", "solve-level-loop-fma",
                              "solve-coop-level-loop", "solve-coop-level-loop-fma", "solve-grid-barrier-level-loop",
                              "solve-grid-barrier-level-loop-fma"])
def test_kernel_variants(ilupp, opts):
    """every schedule / arithmetic variant meets the same parity bar"""
    defaults = {k: ilupp.get_option(k) for k in opts}
    try:
        for k, v in opts.items():
            ilupp.set_option(k, v)
        for A in (P.poisson3d(11), P.irregular_powerlaw(5000, 9, seed=8), P.convdiff3d(10).tocsc()):
            A.indices = A.indices.astype(np.int32)
            A.indptr = A.indptr.astype(np.int32)
            b = P.rhs(A.shape[0], 4)
            B = ilupp.ILU0Preconditioner(A)
            Lo, Uo = oracle.port.ilu0(A)
            L, U = B.factors()
            assert same_pattern(L, Lo) and same_pattern(U, Uo)
            assert relerr(L.data, Lo.data) <= TOL and relerr(U.data, Uo.data) <= TOL
            for tr in (False, True):
                y = b.copy()
                (B.apply_trans if tr else B.apply)(y)
                assert relerr(y, oracle.port.apply_lu(Lo, Uo, b, tr)) <= TOL
    finally:
        for k, v in defaults.items():
            ilupp.set_option(k, v)


# ---- behaviours at the boundary ---------------------------------------------------------------------------
def test_apply_shapes_and_errors(ilupp):
    A = P.poisson2d(8)
    B = ilupp.ILU0Preconditioner(A)
    n = A.shape[0]
    b = P.rhs(n, 0)
    y1 = B @ b
    y2 = (B @ b.reshape(n, 1))
    assert y1.shape == (n,) and y2.shape == (n, 1) and np.array_equal(y1, y2.ravel())
    x = b.reshape(n, 1).copy()
    B.apply(x)                                           # in place through ravel() (ilupp/__init__.py:144-146)
    assert np.array_equal(x.ravel(), y1)
    with pytest.raises(RuntimeError, match="vector has wrong size for preconditioner!"):    # binding.cpp:241-242
        B.apply(np.zeros(n + 1))
    with pytest.raises(RuntimeError, match=r"Expected d \(d\) array for b, got f!"):
        B.apply(np.zeros(n, np.float32))
    import scipy.sparse.linalg as spla
    xs, info = spla.gmres(A, b, M=B, rtol=1e-12, atol=1e-12, restart=30)     # usable as scipy preconditioner
    assert info == 0 and np.linalg.norm(A @ xs - b) < 1e-8 * np.linalg.norm(b)


def test_structure_validation(ilupp):
    # structurally missing diagonal: undefined behaviour in the reference (ILU0.hpp:31,39-44); refused here
    M = scipy.sparse.csr_matrix((np.array([1.0, 2.0, 3.0]), np.array([0, 1, 0], np.int32), np.array([0, 2, 3], np.int32)),
                                shape=(2, 2))
    with pytest.raises(RuntimeError, match="missing diagonal"):
        ilupp._ilupp.ILU0Preconditioner(M.data, M.indices, M.indptr, True)
    # duplicate column indices break the reference's two-pointer merges; refused here
    with pytest.raises(RuntimeError, match="not strictly ascending"):
        ilupp._ilupp.ILU0Preconditioner(np.ones(4), np.array([0, 0, 0, 1], np.int32), np.array([0, 2, 4], np.int32), True)
    with pytest.raises(RuntimeError, match="out of range"):
        ilupp._ilupp.ILU0Preconditioner(np.ones(2), np.array([0, 5], np.int32), np.array([0, 1, 2], np.int32), True)


def test_nan_inf_propagate_like_the_reference(ilupp):
    A = (-P.poisson2d(6)).tocsr()                                   # not SPD: NaN factor, no error
    (L,) = ilupp.IChol0Preconditioner(A).factors()
    Lo = oracle.port.ichol0(A)
    assert np.array_equal(np.isnan(L.data), np.isnan(Lo.data)) and np.isnan(L.data).any()
    Z = scipy.sparse.csr_matrix((np.array([0.0, 1.0, 1.0, 0.0]), np.array([0, 1, 0, 1], np.int32),
                                 np.array([0, 2, 4], np.int32)), shape=(2, 2))
    L, U = ilupp.ilu0(Z)                                            # stored zero pivot: inf, no error
    Lo, Uo = oracle.port.ilu0(Z)
    assert np.array_equal(L.data, Lo.data, equal_nan=True) and np.array_equal(U.data, Uo.data, equal_nan=True)
    # a NaN right-hand side entry propagates and does not hang the sync-free kernels
    A = P.poisson2d(12)
    B = ilupp.ILU0Preconditioner(A)
    b = np.ones(A.shape[0])
    b[5] = np.nan
    y = b.copy()
    B.apply(y)
    Lo, Uo = oracle.port.ilu0(A)
    assert np.array_equal(np.isnan(y), np.isnan(oracle.port.apply_lu(Lo, Uo, b)))


def test_refactor_and_batched(ilupp):
    A = P.poisson2d(16)
    B = ilupp.ILU0Preconditioner(A)
    A2 = A.copy()
    A2.data = A2.data * (1.0 + 0.1 * np.random.default_rng(0).random(A2.nnz))
    B.refactor(A2)
    L, U = B.factors()
    Lo, Uo = oracle.port.ilu0(A2)
    assert np.array_equal(L.data, Lo.data) and np.array_equal(U.data, Uo.data)
    pat, vals = P.batched_poisson2d(5, 12)
    n1 = pat.shape[0]
    Bb = ilupp.ILU0PreconditionerBatched(pat, vals)
    assert Bb.shape == (5 * n1, 5 * n1)
    X = P.rhs(5 * n1, 9)
    y = X.copy()
    Bb.apply(y)
    for s in range(5):
        As = pat.copy()
        As.data = vals[s].copy()
        Lo, Uo = oracle.port.ilu0(As)
        assert np.array_equal(y[s * n1:(s + 1) * n1], oracle.port.apply_lu(Lo, Uo, X[s * n1:(s + 1) * n1]))


def test_torch_device_tensors(ilupp):
    import torch
    A = P.poisson3d(10)
    B = ilupp.IChol0Preconditioner(A)
    b = P.rhs(A.shape[0], 0)
    x = torch.from_numpy(b).cuda()
    B.apply(x)                                          # zero-copy, asynchronous on the current stream
    y = b.copy()
    B.apply(y)
    assert np.array_equal(x.cpu().numpy(), y)
    X = torch.from_numpy(np.asfortranarray(P.rhs(A.shape[0], 1, k=3))).cuda()    # column-major (n, k)
    Xh = X.cpu().numpy().copy()
    B.apply(X.t().contiguous().t())                     # strides (1, n)
    with pytest.raises(RuntimeError):
        B.apply(torch.zeros(7, dtype=torch.float64, device="cuda"))


# ---- BASELINE.json sizes: parity against the oracle + size-independent properties -------------------------
def test_baseline_config1_poisson2d_256(ilupp):
    A = P.poisson2d(256)
    b = P.rhs(A.shape[0], 0)
    for fmt in ("csr", "csc"):
        M = A.asformat(fmt)
        B = ilupp.ILU0Preconditioner(M)
        assert B.total_nnz == 392192
        Lo, Uo = oracle.port.ilu0(M)
        L, U = B.factors()
        assert same_pattern(L, Lo) and same_pattern(U, Uo) and relerr(L.data, Lo.data) <= TOL and relerr(U.data, Uo.data) <= TOL
        y = b.copy()
        B.apply(y)
        assert relerr(y, oracle.port.apply_lu(Lo, Uo, b)) <= TOL
        lev, nlev = B.levels(0)
        assert nlev == 511
        if fmt == "csr":
            i = np.arange(A.shape[0])
            assert np.array_equal(lev, i % 256 + i // 256)            # closed form


def test_baseline_config2_poisson3d_128(ilupp):
    N = 128
    A = P.poisson3d(N)
    n = A.shape[0]
    B = ilupp.IChol0Preconditioner(A)
    assert B.total_nnz == 8339456
    lev, nlev = B.levels(0)
    i = np.arange(n)
    assert nlev == 3 * N - 2 and np.array_equal(lev, i % N + (i // N) % N + i // (N * N))
    (L,) = B.factors()
    Lo = oracle.port.ichol0(A)                                        # ~0.2 s on the host
    assert same_pattern(L, Lo) and relerr(L.data, Lo.data) <= TOL
    b = P.rhs(n, 0)
    y = b.copy()
    B.apply(y)
    assert relerr(y, oracle.port.apply_llt(Lo, b)) <= TOL
    # size-independent properties: L L^T y == b, and linearity of the apply
    r = L @ (L.T @ y)
    assert relerr(r, b) < 1e-10
    b2 = P.rhs(n, 1)
    y2 = b2.copy()
    B.apply(y2)
    y3 = (2.0 * b + b2).copy()
    B.apply(y3)
    assert relerr(y3, 2.0 * y + y2) < 1e-12


def test_baseline_config3_convdiff_bicgstab(ilupp):
    A = P.convdiff3d(48)                                              # reduced size of config 3 (oracle in seconds)
    n = A.shape[0]
    b = A @ np.ones(n)
    x, (it, rel, ab) = ilupp.solve(A, b, rtol=1e-8, atol=1e-8, max_iter=2000, info=True)
    Lo, Uo = oracle.port.ilu0(A)
    xo, ito, relo, abo, ok = oracle.port.bicgstab(A, b, (Lo, Uo), 1e-8, 1e-8, 2000)
    assert ok and abs(it - ito) <= 2
    assert np.abs(x - 1.0).max() < 1e-6 and relerr(x, xo) < 1e-6
    # one preconditioned mat-vec, tight: L^-1 A U^-1 p
    B = ilupp.ILU0Preconditioner(A)
    p = P.rhs(n, 3)
    t = oracle.port.trsv("bwd_gather", Uo, p)
    t = oracle.port.spmv(A, t)
    t = oracle.port.trsv("fwd_gather", Lo, t)
    # GPU: same composition through the public pieces
    L, U = B.factors()
    assert relerr(B.pr.spmv(oracle.port.trsv("bwd_gather", U, p)), oracle.port.spmv(A, oracle.port.trsv("bwd_gather", Uo, p))) <= 1e-13
    with pytest.raises(RuntimeError, match="did not converge"):
        ilupp.solve(A, b, rtol=1e-14, atol=1e-14, max_iter=2)
    with pytest.raises(RuntimeError, match="right-hand side has wrong size!"):
        ilupp.solve(A, b[:-1])


def test_baseline_config4_irregular(ilupp):
    A = P.irregular_powerlaw(60000, 19, seed=1234)                    # reduced n (oracle factor in seconds)
    b = P.rhs(A.shape[0], 0)
    B = ilupp.ILU0Preconditioner(A)
    Lo, Uo = oracle.port.ilu0(A)
    L, U = B.factors()
    assert same_pattern(L, Lo) and same_pattern(U, Uo)
    assert relerr(L.data, Lo.data) <= TOL and relerr(U.data, Uo.data) <= TOL
    y = b.copy()
    B.apply(y)
    assert relerr(y, oracle.port.apply_lu(Lo, Uo, b)) <= TOL
    lev, nlev = B.levels(0)
    lo, nlo = oracle.port.levels(A.indptr, A.indices, lower=True)
    assert nlev == nlo and np.array_equal(lev, lo)


def _level_loop_stress_matrix(n=30000, hub=3000, seed=5):
    """One level wider than the level loop's shared-memory rings, then a deep chain, then hub rows longer than the
    staging limit: exercises the global fall-backs of k_trsv_deep (ilub_deep.cu)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    wide = n - 2000                      # rows [0, wide) have no dependencies at all: one level of `wide` rows
    for i in range(wide, n):             # a chain with a few random far dependencies
        deps = {i - 1} | set(rng.integers(0, i, size=3).tolist())
        if (i - wide) % 400 == 7:        # hub rows: `hub` dependencies
            deps |= set(rng.integers(0, i, size=hub).tolist())
        for j in deps:
            rows.append(i), cols.append(int(j)), vals.append(-rng.random())
    L = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    A = L + L.T
    A = A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)
    A = A.tocsr()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32)
    A.indptr = A.indptr.astype(np.int32)
    return A


@pytest.mark.parametrize("hub", [3000, 5000], ids=["long-rows", "unstaged-blocks"])
@pytest.mark.parametrize("exact", [1, 0])
def test_level_loop_solve_fallbacks(ilupp, exact, hub):
    """wide level (> ring), hub rows (> staging limit, > entry ring) and the deep chain through the level-loop solve"""
    A = _level_loop_stress_matrix(hub=hub)
    b = P.rhs(A.shape[0], 3)
    saved = {k: ilupp.get_option(k) for k in ("solve_algo", "exact")}
    try:
        ilupp.set_option("solve_algo", 3)
        ilupp.set_option("exact", exact)
        B = ilupp.ILU0Preconditioner(A)
        assert B.stats["levels"][0] > 1500
        Lo, Uo = oracle.port.ilu0(A)
        for tr in (False, True):
            y = b.copy()
            (B.apply_trans if tr else B.apply)(y)
            yo = oracle.port.apply_lu(Lo, Uo, b, tr)
            assert relerr(y, yo) <= TOL
            if exact:
                assert np.array_equal(y, yo)
    finally:
        for k, v in saved.items():
            ilupp.set_option(k, v)


def test_handles_reuse_cached_device_memory(ilupp):
    """creating and destroying handles through the caching allocator never changes results"""
    A = P.poisson3d(20)
    b = P.rhs(A.shape[0], 1)
    ref = None
    for _ in range(6):
        B = ilupp.IChol0Preconditioner(A)
        C = ilupp.ILU0Preconditioner(P.convdiff3d(14))
        y = b.copy()
        B.apply(y)
        if ref is None:
            ref = y
        assert np.array_equal(y, ref)
        del B, C


STENCILS = [("poisson2d_33", lambda: P.poisson2d(33), True), ("poisson3d_20", lambda: P.poisson3d(20), True),
            ("convdiff3d_13", lambda: P.convdiff3d(13), False), ("poisson2d_128", lambda: P.poisson2d(128), True)]


@pytest.mark.parametrize("name,gen,sym", STENCILS, ids=[p[0] for p in STENCILS])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_stencil_factorisation(ilupp, name, gen, sym, fmt):
    """fast arithmetic on a stencil matrix: the pivot recurrence on the line-systolic kernel (factor_stencil)"""
    A = gen().asformat(fmt)
    b = P.rhs(A.shape[0], 2)
    saved = ilupp.get_option("exact")
    try:
        ilupp.set_option("exact", 0)
        B = ilupp.ILU0Preconditioner(A)
        Lo, Uo = oracle.port.ilu0(A)
        L, U = B.factors()
        assert same_pattern(L, Lo) and same_pattern(U, Uo)
        assert relerr(L.data, Lo.data) <= TOL and relerr(U.data, Uo.data) <= TOL
        y = b.copy()
        B.apply(y)
        assert relerr(y, oracle.port.apply_lu(Lo, Uo, b)) <= TOL
        A2 = A.copy()
        A2.data = A.data * (1.0 + 0.1 * np.cos(np.arange(A.nnz)))   # numeric re-factorisation on the same pattern
        if sym:
            A2 = ((A2 + A2.T) * 0.5).asformat(fmt)
            A2.sort_indices()
        B.refactor(A2)
        Lo2, Uo2 = oracle.port.ilu0(A2)
        L2, U2 = B.factors()
        assert relerr(L2.data, Lo2.data) <= TOL and relerr(U2.data, Uo2.data) <= TOL
        if sym:
            C = ilupp.IChol0Preconditioner(A)
            (Lc,) = C.factors()
            Lco = oracle.port.ichol0(A)
            assert same_pattern(Lc, Lco) and relerr(Lc.data, Lco.data) <= TOL
            y = b.copy()
            C.apply(y)
            assert relerr(y, oracle.port.apply_llt(Lco, b)) <= TOL
    finally:
        ilupp.set_option("exact", saved)


def test_baseline_config5_batched(ilupp):
    nsys, N = 24, 128
    pat, vals = P.batched_poisson2d(nsys, N)
    n1 = pat.shape[0]
    Bb = ilupp.ILU0PreconditionerBatched(pat, vals)
    assert Bb.total_nnz == nsys * 97792
    X = P.rhs(nsys * n1, 10**6, k=16)
    Y = Bb @ X
    for s in (0, 7, nsys - 1):
        As = pat.copy()
        As.data = vals[s].copy()
        Lo, Uo = oracle.port.ilu0(As)
        for c in (0, 15):
            assert relerr(Y[s * n1:(s + 1) * n1, c], oracle.port.apply_lu(Lo, Uo, X[s * n1:(s + 1) * n1, c])) <= TOL


# ---- plane-per-warp solve (ilub_plane.cu): 2-D grids / batches, many right-hand sides ------------------------
def _grid2d(nx, ny, seed):
    """nonsymmetric 5-point stencil on an nx x ny grid (x fastest), random coefficients, diagonally dominant"""
    rng = np.random.default_rng(seed)
    n = nx * ny
    i = np.arange(n)
    rows, cols = [], []
    for off, ok in ((-1, i % nx > 0), (1, i % nx < nx - 1), (-nx, i >= nx), (nx, i < n - nx)):
        rows.append(i[ok])
        cols.append(i[ok] + off)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    A = scipy.sparse.csr_matrix((-rng.uniform(0.5, 1.5, rows.size), (rows, cols)), shape=(n, n))
    A = (A + scipy.sparse.diags(1.0 + np.asarray(abs(A).sum(axis=1)).ravel())).tocsr()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32)
    A.indptr = A.indptr.astype(np.int32)
    return A


PLANE_GRIDS = [("1-block", 20, 20), ("2-blocks-ragged", 37, 50), ("4-blocks-ragged", 70, 100), ("4-blocks", 128, 128),
               ("8-blocks", 256, 256), ("wide", 300, 9), ("tiny", 3, 5), ("two-columns", 2, 70), ("ragged-8", 9, 230)]


@pytest.mark.parametrize("name,nx,ny", PLANE_GRIDS, ids=[g[0] for g in PLANE_GRIDS])
@pytest.mark.parametrize("exact", [1, 0])
def test_plane_solve_matches_line_systolic_and_oracle(ilupp, name, nx, ny, exact):
    """the plane-per-warp kernel runs the same recurrence in the same order as the line-systolic kernel: results are
    bit-for-bit equal in both arithmetic modes, for every solve form, any number of right-hand sides per warp / CTA"""
    A = _grid2d(nx, ny, 3)
    n = A.shape[0]
    X = np.asfortranarray(P.rhs(n, 5, k=5))
    keys = ("exact", "sys_plane", "plane_r", "plane_wpc")
    defaults = {k: ilupp.get_option(k) for k in keys}
    try:
        ilupp.set_option("exact", exact)
        B = ilupp.ILU0Preconditioner(A)
        assert B.stats["part_mode"] == 4                              # line-systolic layout
        Lo, Uo = oracle.port.ilu0(A)
        ilupp.set_option("sys_plane", 0)
        ref = [B._matmat(X), B._rmatmat(X)]
        launches0 = B.stats["launches"]
        for r, wpc in ((1, 0), (2, 0), (1, 2), (2, 1)):
            ilupp.set_option("sys_plane", 2)
            ilupp.set_option("plane_r", r)
            ilupp.set_option("plane_wpc", wpc)
            got = [B._matmat(X), B._rmatmat(X)]
            for tr in (0, 1):
                assert np.array_equal(got[tr], ref[tr]), (r, wpc, tr)
            y = X[:, 0].copy()
            B.apply(y)                                                # single right-hand side through apply()
            assert np.array_equal(y, ref[0][:, 0])
        assert B.stats["launches"] - launches0 == 4 * 6               # one launch per triangular solve, all columns
        for tr in (False, True):
            for c in (0, 4):
                yo = oracle.port.apply_lu(Lo, Uo, np.ascontiguousarray(X[:, c]), tr)
                assert relerr(ref[int(tr)][:, c], yo) <= TOL
                if exact:
                    assert np.array_equal(ref[int(tr)][:, c], yo)
    finally:
        for k, v in defaults.items():
            ilupp.set_option(k, v)


@pytest.mark.parametrize("exact", [1, 0])
@pytest.mark.parametrize("kind", ["ilu0", "ichol0"])
def test_plane_solve_batched(ilupp, kind, exact):
    nsys, N, k = 7, 40, 19
    pat, vals = P.batched_poisson2d(nsys, N)
    n1 = pat.shape[0]
    keys = ("exact", "sys_plane", "plane_r")
    defaults = {k_: ilupp.get_option(k_) for k_ in keys}
    try:
        ilupp.set_option("exact", exact)
        Bb = (ilupp.ILU0PreconditionerBatched if kind == "ilu0" else ilupp.IChol0PreconditionerBatched)(pat, vals)
        X = np.asfortranarray(P.rhs(nsys * n1, 11, k=k))
        ilupp.set_option("sys_plane", 0)
        ref = [Bb._matmat(X), Bb._rmatmat(X)]
        for r in (1, 2):
            ilupp.set_option("sys_plane", 1)                          # auto: batches take the plane kernel
            ilupp.set_option("plane_r", r)
            l0 = Bb.stats["launches"]
            got = [Bb._matmat(X), Bb._rmatmat(X)]
            assert Bb.stats["launches"] - l0 == 4
            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
        for s in (0, 3, nsys - 1):
            As = pat.copy()
            As.data = vals[s].copy()
            sl = slice(s * n1, (s + 1) * n1)
            for c in (0, k - 1):
                b = np.ascontiguousarray(X[sl, c])
                if kind == "ilu0":
                    Lo, Uo = oracle.port.ilu0(As)
                    yo = oracle.port.apply_lu(Lo, Uo, b)
                else:
                    yo = oracle.port.apply_llt(oracle.port.ichol0(As), b)
                assert relerr(ref[0][sl, c], yo) <= TOL
                if exact:
                    assert np.array_equal(ref[0][sl, c], yo)
    finally:
        for k_, v in defaults.items():
            ilupp.set_option(k_, v)


# ---- GPU apply for factors computed elsewhere (reference ILUT / ILUC / ICholT) --------------------
@pytest.mark.parametrize("name,key", [("laplace2d", "ilut"), ("laplace2d", "iluc"), ("laplace2d", "icholt"),
                                      ("random", "ilut"), ("random", "iluc")])
def test_factors_preconditioner_golden(ilupp, name, key):
    g = load_golden("factors_" + name)
    L = golden_factor(g, key + "_L")
    U = None if key == "icholt" else golden_factor(g, key + "_U")
    B = ilupp.FactorsPreconditioner(L, U)
    b = g["b"]
    y = b.copy()
    B.apply(y)
    assert relerr(y, g[key + "_apply"]) <= TOL
    yt = b.copy()
    B.apply_trans(yt)
    assert relerr(yt, g[key + "_apply_trans"]) <= TOL
    X = np.asfortranarray(P.rhs(b.size, 2, k=3))
    Y = B @ X
    Lc = L.tocsr()
    for c in range(3):
        yo = oracle.port.apply_llt(Lc, X[:, c]) if U is None else oracle.port.apply_lu(Lc, U.tocsr(), X[:, c])
        assert relerr(Y[:, c], yo) <= TOL
    # the handle hands the factors back in the reference's layout
    F = B.factors()
    assert abs(F[0] - L).max() == 0
    if U is not None:
        assert abs(F[1] - U).max() == 0
        assert B.total_nnz == Lc.nnz + U.nnz


def test_factors_preconditioner_runs_the_same_kernels(ilupp):
    """zero-fill factors handed back in: same kernels, same values -> bit-for-bit the same apply (line-systolic,
    level-loop and chunked sync-free paths)"""
    for A in (P.poisson3d(14), P.poisson2d(40), P.irregular_powerlaw(4000, 9, seed=3), P.convdiff3d(9).tocsc()):
        A.indices = A.indices.astype(np.int32)
        A.indptr = A.indptr.astype(np.int32)
        b = P.rhs(A.shape[0], 6)
        B0 = ilupp.ILU0Preconditioner(A)
        L, U = B0.factors()
        B1 = ilupp.FactorsPreconditioner(L, U)
        for tr in (False, True):
            y0, y1 = b.copy(), b.copy()
            (B0.apply_trans if tr else B0.apply)(y0)
            (B1.apply_trans if tr else B1.apply)(y1)
            if A.format == "csr":
                assert np.array_equal(y0, y1)
            assert relerr(y1, oracle.port.apply_lu(L, U, b, tr)) <= TOL
    A = P.poisson3d(12)
    C0 = ilupp.IChol0Preconditioner(A)
    (Lc,) = C0.factors()
    C1 = ilupp.FactorsPreconditioner(Lc)
    b = P.rhs(A.shape[0], 6)
    y0, y1 = b.copy(), b.copy()
    C0.apply(y0)
    C1.apply(y1)
    assert np.array_equal(y0, y1)
    with pytest.raises(ValueError):
        ilupp.FactorsPreconditioner(U, U)                             # not lower / no unit diagonal
    with pytest.raises(ValueError):
        ilupp.FactorsPreconditioner(scipy.sparse.tril(A, -1) + 2 * scipy.sparse.eye(A.shape[0]), scipy.sparse.triu(A))


# ---- BASELINE.json configs at FULL size (round 2: the sizes the metric is quoted on, driver-run) ----------------
def test_full_config2_fast_arithmetic_128(ilupp):
    """config 2 at 128^3 in the arithmetic mode bench.py runs (exact=0: FMA, reciprocal diagonal, stencil factorisation)"""
    N = 128
    A = P.poisson3d(N)
    n = A.shape[0]
    saved = ilupp.get_option("exact")
    try:
        ilupp.set_option("exact", 0)
        B = ilupp.IChol0Preconditioner(A)
        (L,) = B.factors()
        Lo = oracle.port.ichol0(A)
        assert same_pattern(L, Lo) and relerr(L.data, Lo.data) <= TOL
        for seed in (0, 5):
            b = P.rhs(n, seed)
            y = b.copy()
            B.apply(y)
            assert relerr(y, oracle.port.apply_llt(Lo, b)) <= TOL
        # the same through a device tensor (the path bench.py times)
        import torch
        b = P.rhs(n, 7)
        x = torch.from_numpy(b).cuda()
        B.apply(x)
        assert relerr(x.cpu().numpy(), oracle.port.apply_llt(Lo, b)) <= TOL
        B2 = ilupp.ILU0Preconditioner(A)
        Lo2, Uo2 = oracle.port.ilu0(A)
        y = b.copy()
        B2.apply_trans(y)
        assert relerr(y, oracle.port.apply_lu(Lo2, Uo2, b, True)) <= TOL
    finally:
        ilupp.set_option("exact", saved)


REF_BICGSTAB_ITERS_160 = 139   # the unmodified reference (oracle/_ref, ILU0 + bicgstab SPLIT) on this problem, recorded
                               # on the GPU box's host in round 1 (profiles/r1i_all_configs.json: 42 s of CPU time)


@pytest.mark.parametrize("exact", [1, 0])
def test_full_config3_convdiff_160(ilupp, exact):
    """config 3 at 160^3: factors, apply, one preconditioned mat-vec against the oracle; BiCGStab to 1e-8"""
    A = P.convdiff3d(160)
    n = A.shape[0]
    saved = ilupp.get_option("exact")
    try:
        ilupp.set_option("exact", exact)
        B = ilupp.ILU0Preconditioner(A)
        Lo, Uo = oracle.port.ilu0(A)                                  # 0.4 s on the host
        L, U = B.factors()
        assert same_pattern(L, Lo) and same_pattern(U, Uo)
        assert relerr(L.data, Lo.data) <= TOL and relerr(U.data, Uo.data) <= TOL
        if exact:
            assert np.array_equal(L.data, Lo.data) and np.array_equal(U.data, Uo.data)
        b = P.rhs(n, 3)
        for tr in (False, True):
            y = b.copy()
            (B.apply_trans if tr else B.apply)(y)
            yo = oracle.port.apply_lu(Lo, Uo, b, tr)
            assert relerr(y, yo) <= TOL
            if exact:
                assert np.array_equal(y, yo)
        # one preconditioned mat-vec of the SPLIT iteration: L^-1 A U^-1 p, every piece on the GPU
        p = P.rhs(n, 4)
        t_o = oracle.port.trsv("fwd_gather", Lo, oracle.port.spmv(A, oracle.port.trsv("bwd_gather", Uo, p)))
        t_g = B.pr.spmv(oracle.port.trsv("bwd_gather", Uo, p))
        assert relerr(t_g, oracle.port.spmv(A, oracle.port.trsv("bwd_gather", Uo, p))) <= 1e-13
        lev, nlev = B.levels(0)
        i = np.arange(n)
        assert nlev == 3 * 160 - 2 and np.array_equal(lev, i % 160 + (i // 160) % 160 + i // 25600)
        assert np.isfinite(t_o).all()
        if exact:
            # BiCGStab: same algorithm, but the GPU's dot products are tree reductions and its SpMV uses FMA, so the
            # iteration history is not bit-identical; the count stays within 10 % of the reference's recorded run and
            # the solution meets the same tolerance (rounding sensitivity of BiCGStab, not a different method)
            bb = A @ np.ones(n)
            x, (it, rel, ab) = B.solve(bb, rtol=1e-8, atol=1e-8, max_iter=2000, info=True)
            assert abs(it - REF_BICGSTAB_ITERS_160) <= 14, it
            assert np.abs(x - 1.0).max() < 1e-6
            assert rel <= 1e-8 and ab <= 1e-8
    finally:
        ilupp.set_option("exact", saved)


def test_full_config4_irregular_1m(ilupp):
    """config 4 at n = 10^6 (nnz ~ 18.8 M, 113,837 levels): pattern, levels ==, values / apply <= 1e-12"""
    A = P.irregular_powerlaw()
    n = A.shape[0]
    b = P.rhs(n, 0)
    B = ilupp.ILU0Preconditioner(A)
    Lo, Uo = oracle.port.ilu0(A)                                      # 2.5 s on the host
    L, U = B.factors()
    assert same_pattern(L, Lo) and same_pattern(U, Uo)
    assert relerr(L.data, Lo.data) <= TOL and relerr(U.data, Uo.data) <= TOL
    for tr in (False, True):
        y = b.copy()
        (B.apply_trans if tr else B.apply)(y)
        assert relerr(y, oracle.port.apply_lu(Lo, Uo, b, tr)) <= TOL
    lev, nlev = B.levels(0)
    lo, nlo = oracle.port.levels(A.indptr, A.indices, lower=True)
    assert nlev == nlo == 113837 and np.array_equal(lev, lo)
    lev1, nlev1 = B.levels(1)
    lo1, nlo1 = oracle.port.levels(A.indptr, A.indices, lower=False)
    assert nlev1 == nlo1 and np.array_equal(lev1, lo1)


def test_full_config5_batched_4096x16(ilupp):
    """config 5 at full size: 4096 systems x 16 right-hand sides on one GPU, a sample of 64 systems against the oracle"""
    import torch
    nsys, N = 4096, 128
    pat, vals = P.batched_poisson2d(nsys, N)
    n1 = pat.shape[0]
    saved = ilupp.get_option("exact")
    try:
        for exact in (0, 1):
            ilupp.set_option("exact", exact)
            Bb = ilupp.ILU0PreconditionerBatched(pat, vals)
            assert Bb.total_nnz == nsys * 97792
            g = torch.Generator(device="cuda").manual_seed(11)
            X = torch.randn((16, nsys * n1), dtype=torch.float64, device="cuda", generator=g).t()   # column-major (n, 16)
            X0 = X.clone()
            Bb.apply(X)
            torch.cuda.synchronize()
            sample = np.unique(np.concatenate([[0, 1, nsys - 1], np.random.default_rng(3).integers(0, nsys, 64)]))
            assert sample.size >= 60
            for s in sample:
                As = pat.copy()
                As.data = vals[s].copy()
                Lo, Uo = oracle.port.ilu0(As)
                for c in (0, 9, 15):
                    bo = X0[s * n1:(s + 1) * n1, c].cpu().numpy()
                    yo = oracle.port.apply_lu(Lo, Uo, np.ascontiguousarray(bo))
                    yg = X[s * n1:(s + 1) * n1, c].cpu().numpy()
                    assert relerr(yg, yo) <= TOL
                    if exact:
                        assert np.array_equal(yg, yo)
            del Bb, X, X0
    finally:
        ilupp.set_option("exact", saved)


def test_strided_column_view_is_rejected(ilupp):
    """(n, 1) view of a C-ordered (n, k) tensor: stride(0) != 1 must raise, not solve on the wrong elements"""
    import torch
    A = P.poisson2d(16)
    n = A.shape[0]
    B = ilupp.ILU0Preconditioner(A)
    X = torch.ones((n, 4), dtype=torch.float64, device="cuda")
    with pytest.raises(RuntimeError, match="contiguous"):
        B.apply(X[:, :1])
    assert float(X.sum()) == 4.0 * n                                  # untouched
    Xf = torch.ones((4, n), dtype=torch.float64, device="cuda").t()   # column-major: fine
    B.apply(Xf[:, :1])
    y = np.ones(n)
    B.apply(y)
    assert np.array_equal(Xf[:, 0].cpu().numpy(), y)
    B.check_status()


def test_shared_handle_across_threads(ilupp):
    """two Python threads applying one preconditioner (the GIL is released inside apply): serialised by the handle lock"""
    import threading
    A = P.poisson3d(24)
    n = A.shape[0]
    B = ilupp.IChol0Preconditioner(A)
    bs = [P.rhs(n, s) for s in range(4)]
    want = []
    for b in bs:
        y = b.copy()
        B.apply(y)
        want.append(y)
    errs = []

    def work(k):
        for _ in range(20):
            y = bs[k].copy()
            B.apply(y)
            if not np.array_equal(y, want[k]):
                errs.append(k)
    ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs


# ---- tile-systolic kernel (ilub_tile.cu; option sys_tile): 8 x 4 lines per warp, TMA in / out, forwarder warps ------
def _grid3d(nx, ny, nz, seed, sym=False):
    """7-point stencil on an nx x ny x nz grid (x fastest), random coefficients, diagonally dominant"""
    rng = np.random.default_rng(seed)
    n = nx * ny * nz
    i = np.arange(n)
    ix, iy, iz = i % nx, (i // nx) % ny, i // (nx * ny)
    rows, cols = [], []
    for off, ok in ((-1, ix > 0), (1, ix < nx - 1), (-nx, iy > 0), (nx, iy < ny - 1), (-nx * ny, iz > 0), (nx * ny, iz < nz - 1)):
        rows.append(i[ok])
        cols.append(i[ok] + off)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    A = scipy.sparse.csr_matrix((-rng.uniform(0.5, 1.5, rows.size), (rows, cols)), shape=(n, n))
    if sym:
        A = (A + A.T) * 0.5
    A = (A + scipy.sparse.diags(1.0 + np.asarray(abs(A).sum(axis=1)).ravel())).tocsr()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32)
    A.indptr = A.indptr.astype(np.int32)
    return A


TILE_GRIDS = [("one-tile", 16, 8, 4), ("ragged", 10, 9, 5), ("ragged-4-ctas", 30, 21, 13), ("short-lines", 6, 40, 37),
              ("clusters", 64, 50, 30)]


@pytest.mark.parametrize("name,nx,ny,nz", TILE_GRIDS, ids=[g[0] for g in TILE_GRIDS])
@pytest.mark.parametrize("exact", [1, 0])
def test_tile_systolic_kernel(ilupp, exact, name, nx, ny, nz):
    """all four solve forms (apply / apply_trans, csr / csc) through k_trsv_tile; exact mode bit-for-bit, and the stencil
    factorisation on the same kernel in fast arithmetic"""
    saved = {k: ilupp.get_option(k) for k in ("exact", "sys_tile")}
    try:
        ilupp.set_option("exact", exact)
        ilupp.set_option("sys_tile", 1)
        A0 = _grid3d(nx, ny, nz, 7)
        b = P.rhs(A0.shape[0], 1)
        for fmt in ("csr", "csc"):
            A = A0.asformat(fmt)
            B = ilupp.ILU0Preconditioner(A)
            assert B.stats["part_mode"] == 7                          # the tile kernel took it
            Lo, Uo = oracle.port.ilu0(A)
            L, U = B.factors()
            assert relerr(L.data, Lo.data) <= TOL and relerr(U.data, Uo.data) <= TOL
            for tr in (False, True):
                y = b.copy()
                (B.apply_trans if tr else B.apply)(y)
                yo = oracle.port.apply_lu(Lo, Uo, b, tr)
                assert relerr(y, yo) <= TOL
                if exact:
                    assert np.array_equal(y, yo)
        S = _grid3d(nx, ny, nz, 8, sym=True)
        C = ilupp.IChol0Preconditioner(S)
        Lco = oracle.port.ichol0(S)
        (Lc,) = C.factors()
        assert relerr(Lc.data, Lco.data) <= TOL
        y = b.copy()
        C.apply(y)
        assert relerr(y, oracle.port.apply_llt(Lco, b)) <= TOL
    finally:
        for k, v in saved.items():
            ilupp.set_option(k, v)


def test_zero_copy_apply_on_pinned_host_vector(ilupp):
    """apply() on a pinned host vector: the solve kernels read / write it over PCIe (no staging copies); same bits as
    the staged path, for all four solve forms"""
    import torch
    for A, cls in ((P.poisson3d(40), ilupp.IChol0Preconditioner), (P.convdiff3d(33), ilupp.ILU0Preconditioner),
                   (P.poisson2d(96), ilupp.ILU0Preconditioner)):
        n = A.shape[0]
        for fmt in ("csr", "csc"):
            B = cls(A.asformat(fmt))
            b = P.rhs(n, 2)
            for tr in (False, True):
                want = b.copy()
                ilupp.set_option("zero_copy", 0)
                (B.apply_trans if tr else B.apply)(want)
                ilupp.set_option("zero_copy", 1)
                hx = torch.from_numpy(b.copy()).pin_memory()
                launches0 = B.stats["launches"]
                (B.apply_trans if tr else B.apply)(hx.numpy())
                ilupp.set_option("zero_copy", 0)
                assert B.stats["launches"] - launches0 == 2
                assert np.array_equal(hx.numpy(), want)


def test_pcg_with_ichol0(ilupp):
    """preconditioned CG for IChol(0) on the fused Krylov kernels (2 reductions per iteration), against the
    oracle's restatement of the reference's cg; solve(precond='ichol0') and IChol0Preconditioner.solve() use it"""
    for N in (12, 40):
        A = P.poisson3d(N)
        n = A.shape[0]
        b = A @ np.ones(n)
        Lo = oracle.port.ichol0(A)
        xo, ito, relo, abo, ok = oracle.port.cg(A, b, (Lo,), 1e-8, 1e-8, 2000)
        assert ok
        x, (it, rel, ab) = ilupp.solve(A, b, rtol=1e-8, atol=1e-8, max_iter=2000, info=True, precond="ichol0")
        assert abs(it - ito) <= 2 and rel <= 1e-8 and ab <= 1e-8
        assert np.abs(x - 1.0).max() < 1e-6 and relerr(x, xo) < 1e-6
        B = ilupp.IChol0Preconditioner(A)
        x2, (it2, _, _) = B.solve(b, rtol=1e-8, atol=1e-8, max_iter=2000, info=True)           # default: cg
        assert it2 == it and np.array_equal(x2, x)
        x3, (it3, _, _) = B.solve(b, rtol=1e-8, atol=1e-8, max_iter=2000, info=True, method="bicgstab")
        assert np.abs(x3 - 1.0).max() < 1e-6
    with pytest.raises(RuntimeError, match="did not converge"):
        ilupp.solve(A, b, rtol=1e-14, atol=1e-14, max_iter=2, precond="ichol0")


@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("key", ["ilutp", "ilucp"])
def test_permuted_factors_apply(ilupp, fmt, key):
    """SURVEY 8(f)1 remainder: factors of the reference's pivoting preconditioners (ILUTP / ILUCP) applied on the GPU with
    the permutations folded into a gather / scatter around the triangular solves; golden outputs from the reference"""
    g = load_golden("factors_perm_" + fmt)
    L, U = golden_factor(g, key + "_L"), golden_factor(g, key + "_U")
    pl, pr = g[key + "_perm_left"], g[key + "_perm_right"]
    B = ilupp.FactorsPreconditioner(L, U, perm_left=pl if pl.size else None, perm_right=pr if pr.size else None)
    y = g["b"].copy()
    B.apply(y)
    assert relerr(y, g[key + "_apply"]) <= 1e-11
    y = g["b"].copy()
    B.apply_trans(y)
    assert relerr(y, g[key + "_apply_trans"]) <= 1e-11
    import torch
    X = torch.from_numpy(np.asfortranarray(np.stack([g["b"], 2.0 * g["b"]], axis=1))).cuda()   # device path, 2 columns
    Xd = torch.empty((2, X.shape[0]), dtype=torch.float64, device="cuda").t()
    Xd.copy_(X)
    B.apply(Xd)
    assert relerr(Xd[:, 0].cpu().numpy(), g[key + "_apply"]) <= 1e-11
    assert relerr(Xd[:, 1].cpu().numpy(), 2.0 * g[key + "_apply"]) <= 1e-11
    with pytest.raises(ValueError):
        ilupp.FactorsPreconditioner(L, U, perm_left=np.zeros(L.shape[0], dtype=np.int32))


# ---- per-handle options, general-path selection, batched analysis ------------------------------------
def test_per_handle_options_and_threads(ilupp):
    """Options belong to the object (the reference's preconditioner objects carry their own parameters and are
    re-entrant, preconditioner.h:140-260): two objects with different arithmetic modes, driven from two threads at the
    same time, give what each gives alone; the process-wide default is untouched."""
    import threading
    A = P.poisson3d(24)
    b = P.rhs(A.shape[0], 3)
    Lo = oracle.port.ichol0(A)
    yo = oracle.port.apply_llt(Lo, b)
    Be = ilupp.IChol0Preconditioner(A, options={"exact": 1})
    Bf = ilupp.IChol0Preconditioner(A, options={"exact": 0})
    assert Be.get_option("exact") == 1 and Bf.get_option("exact") == 0 and ilupp.get_option("exact") == 1
    ye = b.copy(); Be.apply(ye)
    yf = b.copy(); Bf.apply(yf)
    assert np.array_equal(ye, yo)                      # reference operation order: bit for bit
    assert relerr(yf, yo) <= TOL and not np.array_equal(yf, yo)   # fma + reciprocal: rounding differs
    out = {}

    def work(name, B, ref):
        ok = True
        for _ in range(20):
            y = b.copy()
            B.apply(y)
            ok = ok and np.array_equal(y, ref)
        out[name] = ok

    ts = [threading.Thread(target=work, args=("exact", Be, ye)), threading.Thread(target=work, args=("fast", Bf, yf))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert out == {"exact": True, "fast": True}
    Bf.set_option("exact", 1)                          # takes effect with the next call (coefficients are re-packed)
    y = b.copy(); Bf.apply(y)
    assert relerr(y, yo) <= TOL
    assert ilupp.get_option("exact") == 1


def test_few_wide_levels_route(ilupp):
    """A randomly relabelled grid has a handful of very wide levels and no stencil structure: the solve goes through
    one launch per level (stats part_mode 7), results as the oracle's, levels identical."""
    A = P.poisson3d(72)                     # (the route needs at least 16,384 rows per level on average)
    n = A.shape[0]
    perm = np.random.default_rng(1).permutation(n)
    Ap = A[perm][:, perm].tocsr()
    Ap.sort_indices()
    b = P.rhs(n, 1)
    B = ilupp.ILU0Preconditioner(Ap)
    assert B.stats["part_mode"] == 7, B.stats
    L, U = oracle.port.ilu0(Ap)
    y = b.copy(); B.apply(y)
    assert np.array_equal(y, oracle.port.apply_lu(L, U, b))
    y = b.copy(); B.apply_trans(y)
    assert relerr(y, oracle.port.apply_lu(L, U, b, transpose=True)) <= TOL
    lev, nlev = B.levels(0)
    lo, nlo = oracle.port.levels(Ap.indptr, Ap.indices, lower=True)
    assert nlev == nlo and np.array_equal(lev, lo)


def test_batched_levels_replicated_and_closed_form(ilupp):
    """The systems of a batch share one analysis: levels of every system equal the oracle's levels of one system;
    a complete stencil takes the closed form, an incomplete one (dropped couplings) the general recurrence."""
    pat, vals = P.batched_poisson2d(6, 24)
    B = ilupp.ILU0PreconditionerBatched(pat, vals)
    n1 = pat.shape[0]
    for which in (0, 1):
        lev, nlev = B.levels(which)
        lo, nlo = oracle.port.levels(pat.indptr, pat.indices, lower=(which == 0))
        assert nlev == nlo == 2 * 24 - 1
        assert np.array_equal(lev.reshape(6, n1), np.tile(lo, (6, 1)))
    # drop the x-couplings of one grid line: still every dependency at a stencil offset, no longer a complete stencil
    A = P.poisson2d(24).tolil()
    for i in range(24 * 5 + 1, 24 * 6):
        A[i, i - 1] = 0
        A[i - 1, i] = 0
    A = A.tocsr(); A.eliminate_zeros(); A.sort_indices()
    B = ilupp.ILU0Preconditioner(A)
    L, U = oracle.port.ilu0(A)
    lev, nlev = B.levels(0)
    lo, nlo = oracle.port.levels(A.indptr, A.indices, lower=True)
    assert nlev == nlo and np.array_equal(lev, lo)
    b = P.rhs(A.shape[0], 2)
    y = b.copy(); B.apply(y)
    assert np.array_equal(y, oracle.port.apply_lu(L, U, b))


def test_write_binary_matches_reference_files_and_restores(ilupp, tmp_path):
    """Persistence: the factors a handle writes are byte for byte the files the unmodified reference
    writes for its ILU0 factors of the same matrix (tests/golden/isp_*.isp), and a handle restored from them applies
    like the original."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for fmt in ("csr", "csc"):
        g = load_golden("nonsym_tridiag_%s" % fmt)
        A = golden_matrix(g)
        B = ilupp.ILU0Preconditioner(A)
        B.write_binary(tmp_path / ("f_" + fmt))
        for side in ("L", "R"):
            mine = open(tmp_path / ("f_%s-%s.isp" % (fmt, side)), "rb").read()
            assert mine == open(os.path.join(gold, "isp_ilu0_tridiag_%s-%s.isp" % (fmt, side)), "rb").read()
        R = ilupp.FactorsPreconditioner.read_binary(tmp_path / ("f_" + fmt))
        for tr in (False, True):
            y = g["b"].copy(); (B.apply_trans if tr else B.apply)(y)
            z = g["b"].copy(); (R.apply_trans if tr else R.apply)(z)
            assert np.array_equal(y, z)
    S = golden_matrix(load_golden("laplace2d_csr"), "S")
    B = ilupp.IChol0Preconditioner(S)
    B.write_binary(tmp_path / "chol")
    assert not os.path.exists(tmp_path / "chol-R.isp")
    R = ilupp.FactorsPreconditioner.read_binary(tmp_path / "chol")
    b = P.rhs(S.shape[0], 5)
    y = b.copy(); B.apply(y)
    z = b.copy(); R.apply(z)
    assert np.array_equal(y, z) and R.total_nnz == B.total_nnz


def test_block_sequential_solve(ilupp):
    """Deep, narrow, band-local DAGs in the fast arithmetic mode go through the block-sequential kernel (ilub_block.cu,
    stats part_mode 8): all four solves against the oracle, hub rows (cooperative long-row path) and a matrix whose blocks
    have deep local chains included; the level loop (option deep_block = 0) and the exact mode give the same answer."""
    import scipy.sparse as sp
    cases = [P.irregular_powerlaw(n=60000, max_len=2048, seed=7)]
    # a banded matrix with every sub-diagonal inside a 32-row block present: local chains as deep as the block
    n = 20000
    rng = np.random.default_rng(3)
    Bd = sp.diags([rng.uniform(-0.4, -0.1, n - k) for k in (1, 2, 5, 40, 700)], [-1, -2, -5, -40, -700], shape=(n, n))
    cases.append((Bd + Bd.T * 0.5 + sp.eye(n) * 4.0).tocsr())
    for A in cases:
        A.sort_indices()
        A.indices = A.indices.astype(np.int32); A.indptr = A.indptr.astype(np.int32)
        n = A.shape[0]
        b = P.rhs(n, 4)
        Lo, Uo = oracle.port.ilu0(A)
        Bf = ilupp.ILU0Preconditioner(A, options={"exact": 0})
        assert Bf.stats["part_mode"] == 8, Bf.stats
        Bl = ilupp.ILU0Preconditioner(A, options={"exact": 0, "deep_block": 0})
        assert Bl.stats["part_mode"] == 5
        for tr in (False, True):
            yo = oracle.port.apply_lu(Lo, Uo, b, tr)
            y = b.copy(); (Bf.apply_trans if tr else Bf.apply)(y)
            z = b.copy(); (Bl.apply_trans if tr else Bl.apply)(z)
            assert relerr(y, yo) <= TOL and relerr(z, yo) <= TOL
        X = P.rhs(n, 9, k=3)          # several right-hand sides: one CTA per column, the same numbers as one at a time
        Y = Bf @ X
        for c in range(3):
            yc = X[:, c].copy(); Bf.apply(yc)
            assert np.array_equal(Y[:, c], yc)
        Bf.set_option("exact", 1)     # the exact mode keeps the level loop (the factors stay the fast-mode ones)
        y = b.copy(); Bf.apply(y)
        assert relerr(y, oracle.port.apply_lu(Lo, Uo, b)) <= TOL
        Be = ilupp.ILU0Preconditioner(A, options={"exact": 1})
        assert Be.stats["part_mode"] == 5
        y = b.copy(); Be.apply(y)
        assert np.array_equal(y, oracle.port.apply_lu(Lo, Uo, b))


@pytest.mark.parametrize("exact", [1, 0])
def test_block_paths_ragged_size(ilupp, exact, capfd):
    """Deep DAG whose size is not a multiple of the 32-row blocks and large enough for the probe: levels through
    k_levels_block, factorisation through k_factor_block (bit for bit in exact mode), solve through k_trsv_block (fast
    mode) -- all against the oracle."""
    n = 300007
    A = P.irregular_powerlaw(n=n, seed=11)
    b = P.rhs(n, 6)
    B = ilupp.ILU0Preconditioner(A, options={"exact": exact})
    assert B.stats["part_mode"] == (5 if exact else 8), B.stats
    Lo, Uo = oracle.port.ilu0(A)
    L, U = B.factors()
    assert same_pattern(L, Lo) and same_pattern(U, Uo)
    assert relerr(L.data, Lo.data) <= TOL and relerr(U.data, Uo.data) <= TOL
    if exact:
        assert np.array_equal(L.data, Lo.data) and np.array_equal(U.data, Uo.data)
    for which, lower in ((0, True), (1, False)):
        lev, nlev = B.levels(which)
        lo, nlo = oracle.port.levels(A.indptr, A.indices, lower=lower)
        assert nlev == nlo and np.array_equal(lev, lo)
    for tr in (False, True):
        y = b.copy(); (B.apply_trans if tr else B.apply)(y)
        yo = oracle.port.apply_lu(Lo, Uo, b, tr)
        assert relerr(y, yo) <= TOL
        if exact:
            assert np.array_equal(y, yo)
    import torch
    torch.cuda.synchronize()
    assert capfd.readouterr().out == ""   # (the kernels' debug timers print only under option trace)
    # numeric re-factorisation on the unchanged pattern (values perturbed entry by entry: the solve's packed coefficients
    # and the blocks' solved local systems must follow)
    A2 = A.copy(); A2.data = A.data * np.random.default_rng(2).uniform(0.95, 1.05, A.nnz)
    B.refactor(A2)
    L2o, U2o = oracle.port.ilu0(A2)
    L2, U2 = B.factors()
    assert relerr(L2.data, L2o.data) <= TOL and relerr(U2.data, U2o.data) <= TOL
    for tr in (False, True):
        y = b.copy(); (B.apply_trans if tr else B.apply)(y)
        assert relerr(y, oracle.port.apply_lu(L2o, U2o, b, tr)) <= TOL


def test_syncfree_configuration_that_cannot_be_resident_is_refused(ilupp):
    """The chunked sync-free kernel needs every CTA resident.  A configuration that cannot be (eight CTAs per SM do not fit
    the shared memory its chunks need) is refused when the solve is analysed -- no dead-lock, no watchdog; the default
    configuration of the same matrix gives the oracle's results."""
    A = random_matrix(20000, "csr").tocsr()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32); A.indptr = A.indptr.astype(np.int32)
    b = P.rhs(A.shape[0], 8)
    with pytest.raises(RuntimeError):
        ilupp.ILU0Preconditioner(A, options={"solve_algo": 2, "ctas_per_sm": 8, "nparts": 600, "exact": 1})
    B = ilupp.ILU0Preconditioner(A, options={"solve_algo": 2, "exact": 1})
    Lo, Uo = oracle.port.ilu0(A)
    for tr in (False, True):
        y = b.copy(); (B.apply_trans if tr else B.apply)(y)
        assert np.array_equal(y, oracle.port.apply_lu(Lo, Uo, b, tr))
