"""dev helper: random 7-point stencil matrices on ragged grids"""
import numpy as np


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
