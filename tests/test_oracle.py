"""CPU: pin the oracle (oracle/ilub_oracle.c) against
  (1) the golden fixtures generated from the unmodified reference (tests/golden, scripts/make_golden.py),
  (2) the reference's own known-answer tests for this path (test/tests.py:197-207,229-259,301-306,324-338),
  (3) the compiled reference itself (oracle/_ref) when present, bit for bit, on larger seeded inputs,
  (4) closed-form level sets.
No GPU, no product code."""
import numpy as np
import pytest
import scipy.sparse

import oracle
from ilupp_b200 import problems as P
from helpers import (example_laplace, golden_factor, golden_matrix, ichol0_dense, ilu0_dense, laplace2d_matrix,
                     laplace_matrix, load_golden, random_matrix, relerr, same_pattern)

CASES = [(p, f) for p in ("laplace", "laplace2d", "random", "nonsym_tridiag") for f in ("csr", "csc")]


@pytest.mark.parametrize("prob,fmt", CASES)
def test_port_matches_golden_ilu0(prob, fmt):
    g = load_golden("%s_%s" % (prob, fmt))
    A = golden_matrix(g)
    L, U = oracle.port.ilu0(A)
    Lg, Ug = golden_factor(g, "ilu0_L"), golden_factor(g, "ilu0_U")
    assert same_pattern(L, Lg) and same_pattern(U, Ug)
    assert np.array_equal(L.data, Lg.data) and np.array_equal(U.data, Ug.data)      # bit-exact
    assert L.nnz + U.nnz == int(g["ilu0_total_nnz"])
    assert np.array_equal(oracle.port.apply_lu(L, U, g["b"], False), g["ilu0_apply"])
    assert np.array_equal(oracle.port.apply_lu(L, U, g["b"], True), g["ilu0_apply_trans"])


@pytest.mark.parametrize("prob,fmt", CASES)
def test_port_matches_golden_ichol0(prob, fmt):
    g = load_golden("%s_%s" % (prob, fmt))
    S = golden_matrix(g, "S")
    L = oracle.port.ichol0(S)
    Lg = golden_factor(g, "ichol0_L")
    assert L.format == "csr" and same_pattern(L, Lg)                                 # IChol.hpp:67-68: always CSR
    assert np.array_equal(L.data, Lg.data, equal_nan=True)
    assert L.nnz == int(g["ichol0_total_nnz"])
    assert np.array_equal(oracle.port.apply_llt(L, g["b"]), g["ichol0_apply"], equal_nan=True)
    assert np.array_equal(g["ichol0_apply"], g["ichol0_apply_trans"], equal_nan=True)  # LL^T: apply == apply_trans


def test_port_bicgstab_matches_golden():
    g = load_golden("bicgstab_convdiff12")
    n = g["A_indptr"].size - 1
    A = scipy.sparse.csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=(n, n))
    L, U = oracle.port.ilu0(A)
    x, it, rel, ab, ok = oracle.port.bicgstab(A, g["b"], (L, U), 1e-8, 1e-8, 2000)
    assert ok == bool(g["ok"]) and it == int(g["iters"])
    assert np.array_equal(x, g["x"]) and rel == float(g["rel"]) and ab == float(g["abs"])


# ---- the reference's own tests, run against the port ------------------------------------------------
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_solve_in_one_step(fmt):  # tests.py:197-207,286-299: zero-fill factors of a tridiagonal matrix are exact
    A, b, x_exact = example_laplace(50, fmt)
    L, U = oracle.port.ilu0(A)
    assert np.allclose(oracle.port.apply_lu(L, U, b), x_exact)
    assert np.allclose(oracle.port.apply_lu(L, U, A.T @ x_exact, True), x_exact)
    Lc = oracle.port.ichol0(A)
    assert np.allclose(oracle.port.apply_llt(Lc, b), x_exact)


@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_factors_correct_and_total_nnz(fmt):  # tests.py:229-259,301-306
    A = laplace_matrix(50, fmt)
    L, U = oracle.port.ilu0(A)
    assert np.allclose((L @ U).toarray(), A.toarray())
    assert all(i >= j for i, j in zip(*L.nonzero())) and all(i <= j for i, j in zip(*U.nonzero()))
    assert L.nnz + U.nnz == 2 * (2 * 50 - 1)
    Lc = oracle.port.ichol0(A)
    assert np.allclose((Lc @ Lc.T).toarray(), A.toarray()) and Lc.nnz == 2 * 50 - 1


@pytest.mark.parametrize("gen", [laplace_matrix, laplace2d_matrix, random_matrix])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_standalone_factorisations_vs_dense(gen, fmt):  # tests.py:324-338
    A = gen(50, fmt)
    X, Xr = oracle.port.ilu0(A), ilu0_dense(A)
    assert all(np.allclose(Xi.toarray(), Ri) for Xi, Ri in zip(X, Xr))
    S = ((A + A.T) / 2).asformat(fmt)
    assert np.allclose(oracle.port.ichol0(S).toarray(), ichol0_dense(S))


# ---- against the compiled reference, larger seeded inputs, bit for bit ---------------------------------
needs_ref = pytest.mark.skipif(oracle.ref is None, reason="oracle/_ref not built on this machine")


@needs_ref
@pytest.mark.parametrize("name,A,sym", [
    ("poisson2d_24", P.poisson2d(24), True), ("poisson3d_9", P.poisson3d(9), True),
    ("convdiff3d_9", P.convdiff3d(9), False), ("powerlaw_4000", P.irregular_powerlaw(4000, 8, seed=5), False)])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_port_equals_compiled_reference(name, A, sym, fmt):
    M = A.asformat(fmt)
    b = P.rhs(M.shape[0], 1)
    R = oracle.ref.ILU0(M)
    L, U = oracle.port.ilu0(M)
    RL, RU = R.factors()
    assert same_pattern(L, RL) and same_pattern(U, RU)
    assert np.array_equal(L.data, RL.data) and np.array_equal(U.data, RU.data)
    assert R.total_nnz == L.nnz + U.nnz
    for tr in (False, True):
        assert np.array_equal(oracle.port.apply_lu(L, U, b, tr), R.apply(b, tr))
    assert np.array_equal(oracle.port.spmv(M, b), oracle.ref.spmv(M, b))
    bb = M @ np.ones(M.shape[0])
    xp = oracle.port.bicgstab(M, bb, (L, U), 1e-8, 1e-8, 500)
    xr = R.bicgstab(M, bb, 1e-8, 1e-8, 500)
    assert np.array_equal(xp[0], xr[0]) and xp[1:] == xr[1:]
    if sym:
        Rc = oracle.ref.IChol0(M)
        Lc = oracle.port.ichol0(M)
        (RLc,) = Rc.factors()
        assert same_pattern(Lc, RLc) and np.array_equal(Lc.data, RLc.data)
        assert np.array_equal(oracle.port.apply_llt(Lc, b), Rc.apply(b))


@needs_ref
def test_reference_quirks_restated():
    # IChol0 of a non-SPD matrix: NaN factor, no error
    A = (-P.poisson2d(6)).tocsr()
    Lr, = oracle.ref.IChol0(A).factors()
    Lp = oracle.port.ichol0(A)
    assert np.array_equal(np.isnan(Lr.data), np.isnan(Lp.data)) and np.isnan(Lp.data).any()
    # ILU0 with a stored zero pivot: inf/NaN propagate silently
    Z = scipy.sparse.csr_matrix(np.array([[0.0, 1.0], [1.0, 0.0]]))
    Z = scipy.sparse.csr_matrix((np.array([0.0, 1.0, 1.0, 0.0]), np.array([0, 1, 0, 1], np.int32),
                                 np.array([0, 2, 4], np.int32)), shape=(2, 2))
    Lr, Ur = oracle.ref.ILU0(Z).factors()
    Lp, Up = oracle.port.ilu0(Z)
    assert np.array_equal(Lr.data, Lp.data, equal_nan=True) and np.array_equal(Ur.data, Up.data, equal_nan=True)


# ---- level sets: closed forms (no reference counterpart) ----------------------------------------------
def test_levels_closed_forms():
    T = P.poisson1d(40)
    lev, nlev = oracle.port.levels(T.indptr, T.indices, lower=True)
    assert nlev == 40 and np.array_equal(lev, np.arange(40))
    N = 12
    A = P.poisson2d(N)
    lev, nlev = oracle.port.levels(A.indptr, A.indices, lower=True)
    ix, iy = np.arange(N * N) % N, np.arange(N * N) // N
    assert nlev == 2 * N - 1 and np.array_equal(lev, ix + iy)
    lev, nlev = oracle.port.levels(A.indptr, A.indices, lower=False)
    assert nlev == 2 * N - 1 and np.array_equal(lev, (N - 1 - ix) + (N - 1 - iy))
    N = 7
    A = P.poisson3d(N)
    lev, nlev = oracle.port.levels(A.indptr, A.indices, lower=True)
    i = np.arange(N ** 3)
    assert nlev == 3 * N - 2 and np.array_equal(lev, i % N + (i // N) % N + i // (N * N))
    # transposed DAGs of a structurally symmetric matrix equal the plain ones
    lt, nt = oracle.port.levels(A.indptr, A.indices, lower=True, transposed=True)
    assert nt == nlev and np.array_equal(lt, lev)


def test_generators_match_survey_sizes():
    A = P.poisson2d(256)
    assert A.shape[0] == 65536 and A.nnz == 326656
    A = P.poisson3d(32)
    assert A.nnz == 7 * 32 ** 3 - 6 * 32 ** 2
    A = P.convdiff3d(8)
    assert (A != A.T).nnz > 0 and np.all(A.diagonal() > 0)
    pat, vals = P.batched_poisson2d(3, 8)
    assert vals.shape == (3, pat.nnz) and not np.array_equal(vals[0], vals[1])


# ---- the triangular solves on the factors of the reference's other preconditioners ----------
@pytest.mark.parametrize("name,key", [("laplace2d", "ilut"), ("laplace2d", "iluc"), ("laplace2d", "icholt"),
                                      ("random", "ilut"), ("random", "iluc")])
def test_port_apply_on_reference_threshold_factors(name, key):
    """golden fixtures: factors and apply / apply_trans outputs of the reference's ILUT / ILUC / ICholT"""
    from helpers import golden_factor, load_golden, relerr
    g = load_golden("factors_" + name)
    L = golden_factor(g, key + "_L").tocsr()
    if key == "icholt":
        assert relerr(oracle.port.apply_llt(L, g["b"]), g[key + "_apply"]) <= 1e-12
        assert relerr(oracle.port.apply_llt(L, g["b"]), g[key + "_apply_trans"]) <= 1e-12
    else:
        U = golden_factor(g, key + "_U").tocsr()
        assert relerr(oracle.port.apply_lu(L, U, g["b"]), g[key + "_apply"]) <= 1e-12
        assert relerr(oracle.port.apply_lu(L, U, g["b"], True), g[key + "_apply_trans"]) <= 1e-12


def test_port_cg_matches_reference_cg():
    """cg (iterative_solvers_implementation.h:133-216): same iteration count and the same residual history (achieved
    tolerances bit for bit) as the unmodified reference.  The reference's own x is NOT compared: with a zero start it
    accumulates into an uninitialised vector y (:160 `vector_type y(problem_size)` + sparse_implementation.h:98-102
    "doesn't initialize by default"), so its solution carries whatever the allocator returned; the port starts y at 0."""
    if oracle.ref is None:
        pytest.skip("oracle/_ref not built on this machine")
    for N in (8, 14):
        A = P.poisson3d(N)
        n = A.shape[0]
        b = A @ np.ones(n)
        L = oracle.port.ichol0(A)
        x, it, rel, ab, ok = oracle.port.cg(A, b, (L,), 1e-8, 1e-8, 500)
        R = oracle.ref.IChol0(A)
        _, itr, relr, abr, okr = R.cg(A, b, 1e-8, 1e-8, 500)
        assert ok and okr and it == itr and rel == relr and ab == abr
        assert np.abs(x - 1.0).max() < 1e-7


@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("key", ["ilutp", "ilucp"])
def test_permuted_factor_semantics_on_golden(fmt, key):
    """triangular_solve_perm (sparse_implementation.h:4166-4258) restated as gather / scatter around ordinary triangular
    solves: L[perm_left, :] and U[:, perm_right] are triangular, apply = scatter_right(U^-1 L^-1 gather_left(b)); pinned
    on factors, permutations and apply / apply_trans outputs of the reference's ILUTP / ILUCP (pivoting active)"""
    from helpers import golden_factor, load_golden, relerr
    g = load_golden("factors_perm_" + fmt)
    L, U = golden_factor(g, key + "_L").toarray(), golden_factor(g, key + "_U").toarray()
    pl, pr = g[key + "_perm_left"], g[key + "_perm_right"]
    n = L.shape[0]
    assert (pl.size + pr.size) == n and ((pl.size and (pl != np.arange(n)).any()) or (pr.size and (pr != np.arange(n)).any()))
    Lt = L[pl, :] if pl.size else L
    Ut = U[:, pr] if pr.size else U
    assert np.abs(np.triu(Lt, 1)).max() == 0 and np.abs(np.tril(Ut, -1)).max() == 0
    b = g["b"]
    t = b[pl] if pl.size else b
    z = np.linalg.solve(Ut, np.linalg.solve(Lt, t))
    x = np.empty(n)
    if pr.size:
        x[pr] = z
    else:
        x = z
    assert relerr(x, g[key + "_apply"]) <= 1e-10
    t = b[pr] if pr.size else b                      # transposed operator: gather by perm_right, scatter by perm_left
    z = np.linalg.solve(Lt.T, np.linalg.solve(Ut.T, t))
    xt = np.empty(n)
    if pl.size:
        xt[pl] = z
    else:
        xt = z
    assert relerr(xt, g[key + "_apply_trans"]) <= 1e-10
