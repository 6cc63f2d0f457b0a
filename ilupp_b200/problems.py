"""Synthetic inputs for the BASELINE.json configs.

All matrices are scipy CSR with fp64 values, int32 sorted indices and lexicographic grid order
(x fastest).  They are built directly with numpy (no kron) so the 128^3 / 160^3 cases take
well under a second.  Nothing here touches the GPU.
"""
import numpy as np
import scipy.sparse as sp

__all__ = ["poisson1d", "poisson2d", "poisson3d", "convdiff3d", "irregular_powerlaw", "batched_poisson2d",
           "stencil_grid", "rhs"]


def stencil_grid(dims, offsets, coeffs, diag):
    """CSR matrix of a constant-coefficient stencil on a lexicographic grid (Dirichlet cut-off).

    dims: grid extents, fastest axis first.  offsets: list of (axis, +-1).  coeffs: value per offset.
    """
    dims = tuple(int(d) for d in dims)
    n = int(np.prod(dims))
    strides = np.cumprod((1,) + dims[:-1])
    lin = np.arange(n, dtype=np.int64)
    coords = [(lin // strides[a]) % dims[a] for a in range(len(dims))]
    rows = [lin]
    cols = [lin]
    vals = [np.full(n, float(diag))]
    for (axis, step), c in zip(offsets, coeffs):
        ok = (coords[axis] + step >= 0) & (coords[axis] + step < dims[axis])
        r = lin[ok]
        rows.append(r)
        cols.append(r + step * strides[axis])
        vals.append(np.full(r.size, float(c)))
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    vals = np.concatenate(vals)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr)
    A = sp.csr_matrix((vals, cols.astype(np.int32), indptr.astype(np.int32)), shape=(n, n))
    A.has_sorted_indices = True
    return A


def poisson1d(n):
    return stencil_grid((n,), [(0, -1), (0, 1)], [-1.0, -1.0], 2.0)


def poisson2d(N):
    """2-D 5-point Poisson, diag 4 / off -1 (config 1: N=256; config 5: N=128)."""
    return stencil_grid((N, N), [(0, -1), (0, 1), (1, -1), (1, 1)], [-1.0] * 4, 4.0)


def poisson3d(N):
    """3-D 7-point Poisson, diag 6 / off -1 (config 2: N=128)."""
    offs = [(a, s) for a in range(3) for s in (-1, 1)]
    return stencil_grid((N, N, N), offs, [-1.0] * 6, 6.0)


def convdiff3d(N, beta=(20.0, 10.0, 5.0)):
    """First-order upwind convection-diffusion scaled by h^2 (config 3: N=160).

    lower neighbours -1 - beta_d*h, upper neighbours -1, diag 6 + h*sum(beta): nonsymmetric M-matrix.
    """
    h = 1.0 / (N + 1)
    offs, cs = [], []
    for a in range(3):
        offs += [(a, -1), (a, 1)]
        cs += [-1.0 - beta[a] * h, -1.0]
    return stencil_grid((N, N, N), offs, cs, 6.0 + h * sum(beta))


def irregular_powerlaw(n=1_000_000, mean_offdiag=19, alpha=2.1, max_len=4096, band=64, p_band=0.5, seed=1234):
    """Random power-law row lengths, half banded / half uniform columns, strictly diagonally
    dominant (config 4).  Deterministic for a given seed."""
    rng = np.random.default_rng(seed)
    lens = np.floor((rng.pareto(alpha, n) + 1.0) * mean_offdiag * (alpha - 1.0) / alpha).astype(np.int64)
    lens = np.clip(lens, 1, max_len)
    tot = int(lens.sum())
    rows = np.repeat(np.arange(n, dtype=np.int64), lens)
    banded = rng.random(tot) < p_band
    off = rng.integers(1, band + 1, tot) * rng.choice(np.array([-1, 1]), tot)
    cols = np.where(banded, np.clip(rows + off, 0, n - 1), rng.integers(0, n, tot))
    vals = -rng.random(tot)
    keep = cols != rows
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()  # duplicates summed
    A.sort_indices()
    d = 1.0 + np.asarray(abs(A).sum(axis=1)).ravel()
    A = (A + sp.diags(d)).tocsr()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32, copy=False)
    A.indptr = A.indptr.astype(np.int32, copy=False)
    return A


def batched_poisson2d(nsys, N=128, first=0):
    """Config 5: systems first..first+nsys-1, each poisson2d(N) + diag(U(0,1; seed=s)).

    Returns (pattern CSR with the shared indptr/indices, values[nsys, nnz])."""
    A = poisson2d(N)
    n = A.shape[0]
    diag_pos = np.flatnonzero(A.indices == np.repeat(np.arange(n), np.diff(A.indptr)))
    vals = np.tile(A.data, (nsys, 1))
    for k in range(nsys):
        vals[k, diag_pos] += np.random.default_rng(first + k).random(n)
    return A, vals


def rhs(n, seed=0, k=None):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(n) if k is None else rng.standard_normal((n, k))
