"""Shared helpers for the test-suite (test infrastructure: may import oracle/)."""
import os

import numpy as np
import scipy.sparse

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relerr(a, b):
    """max-norm relative error, the parity measure of BASELINE.json (tolerance 1e-12 for fp64 values)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def same_pattern(X, Y):
    return (X.format == Y.format and X.shape == Y.shape and np.array_equal(X.indptr, Y.indptr)
            and np.array_equal(X.indices, Y.indices))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_matrix(g, prefix="A"):
    cls = scipy.sparse.csr_matrix if bool(g["is_csr"]) else scipy.sparse.csc_matrix
    n = g[prefix + "_indptr"].size - 1
    M = cls((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=(n, n))
    M.has_sorted_indices = True
    return M


def golden_factor(g, prefix):
    cls = scipy.sparse.csr_matrix if bool(g[prefix + "_is_csr"]) else scipy.sparse.csc_matrix
    n = g[prefix + "_indptr"].size - 1
    M = cls((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=(n, n))
    M.has_sorted_indices = True
    return M


# the reference's own test problems (test/tests.py:9-43)
def laplace_matrix(n, format='csr'):
    h = 1.0 / (n + 1)
    d = np.ones(n) / (h ** 2)
    return scipy.sparse.diags((-d[:-1], 2 * d, -d[:-1]), (-1, 0, 1)).asformat(format)


def laplace2d_matrix(n_total, format='csr'):
    n = int(np.sqrt(n_total))
    A, I = laplace_matrix(n), scipy.sparse.eye(n)
    return (scipy.sparse.kron(A, I) + scipy.sparse.kron(I, A)).asformat(format)


def random_matrix(n, format='csr', eye_factor=10.0):
    density = min(1.0, 5 / n)
    return (scipy.sparse.random(n, n, density=density, random_state=39273) +
            eye_factor * scipy.sparse.eye(n)).asformat(format)


def example_laplace(n, format='csr'):
    A = laplace_matrix(n, format=format)
    b = np.ones(n)
    X = np.linspace(0, 1, n + 2)[1:-1]
    return A, b, X * (1 - X) / 2


# dense re-implementations the reference's tests use as oracles (test/tests.py:67-82,108-128)
def ichol0_dense(A):
    A = A.toarray()
    L = np.tril(A)
    n = A.shape[0]
    for k in range(n):
        L[k, k] = np.sqrt(L[k, k])
        for i in range(k + 1, n):
            if L[i, k] != 0:
                L[i, k] /= L[k, k]
        for j in range(k + 1, n):
            for i in range(j, n):
                if L[i, j] != 0:
                    L[i, j] -= L[i, k] * L[j, k]
    return L


def ilu0_dense(A):
    is_csc = isinstance(A, scipy.sparse.csc_matrix)
    if is_csc:
        A = A.T
    A = A.toarray()
    n = A.shape[0]
    for i in range(n):
        for k in range(i):
            if A[i, k] != 0:
                l_ik = A[i, k] / A[k, k]
                A[i, k] = l_ik
                for j in range(k + 1, n):
                    if A[i, j] != 0:
                        A[i, j] -= l_ik * A[k, j]
    L, U = np.tril(A, -1), np.triu(A)
    np.fill_diagonal(L, 1.0)
    if is_csc:
        return (U.T, L.T)
    return (L, U)
