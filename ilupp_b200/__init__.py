"""ilupp_b200 -- B200-native drop-in for the zero-fill hot path of c-f-h/ilupp.

Same public names, argument meaning and error behaviour as the reference's ``ilupp/__init__.py``
for the path SURVEY.md section 8 scopes:

    ILU0Preconditioner, IChol0Preconditioner   (reference ilupp/__init__.py:272-293)
    ilu0, ichol0                               (:314-324)
    solve                                      (:85-119; BiCGStab, here with an ILU(0) preconditioner)
    _BaseWrapper: LinearOperator protocol, apply / apply_trans / total_nnz / factors (:122-168)

All arithmetic runs in hand-written CUDA kernels (sm_100a) inside ``libilub.so`` behind the C ABI
of ``include/ilub.h``; the pybind11 module ``_ilupp`` next to this file is the C++ host code over
that ABI.  There is no CPU fallback: importing this package without the built extension raises,
and constructing a preconditioner without a CUDA device raises ``RuntimeError``.

Extensions over the reference (all optional): ``device=`` to pick a GPU, native multi-RHS
``B @ X``, ``apply``/``apply_trans`` on torch CUDA tensors (zero copy), ``refactor`` with new
values on the same pattern, ``levels``/``level_sets`` (parity hooks), batched constructors.
"""

__version__ = '1.0.2+b200.1'

import os as _os

import numpy as np
import scipy.sparse
import scipy.sparse.linalg

try:
    from . import _ilupp
except ImportError as _e:  # fail loudly: there is no pure-Python or CPU path
    raise ImportError(
        "ilupp_b200: the CUDA extension is not built (%s). Run `python -c \"import __graft_entry__ as g; "
        "g.build()\"` or `make -C ilupp_b200/csrc` first; there is no CPU fallback." % (_e,)) from _e
from ._ilupp import iluplusplus_precond_parameter  # noqa: F401  (signature compatibility of solve())

_index_size = _ilupp.index_size()
if _index_size == 4:
    _index_dtype = np.dtype(np.int32)
elif _index_size == 8:
    _index_dtype = np.dtype(np.int64)
else:
    raise RuntimeError('invalid index type size %d' % _index_size)


def set_option(key, value):
    """Set a global libilub option (see include/ilub.h, ilub_set_option)."""
    _ilupp.set_option(key, int(value))


def get_option(key):
    return _ilupp.get_option(key)


def _stage_options(options):
    """Options for the object about to be created (thread-local in the library, taken over by the next creation on this
    thread): they are the new object's own settings from its first kernel on, whatever the process-wide defaults are."""
    if options:
        for k, v in options.items():
            _ilupp.stage_option(k, int(v))


def _upcast_indices(idx):
    """reference ilupp/__init__.py:38-53: make the index array match the compiled index width."""
    sz, target_sz = idx.dtype.itemsize, _index_size
    if sz == target_sz:
        return idx
    elif sz < target_sz:
        return idx.astype(_index_dtype)
    else:
        raise TypeError(
            'Index array has %d bytes per index, but the library '
            'is compiled for %d bytes per index. Downcasting might '
            'lead to integer overflow. Please compile ilupp with a '
            'larger index type if you need to use very large matrices.'
            % (sz, target_sz))


def _matrix_fields(A):
    """reference ilupp/__init__.py:55-71: (data, indices, indptr, is_csr) of a square scipy CSR/CSC
    matrix; sorts the caller's matrix in place exactly like the reference does.  scipy's newer
    csr_array / csc_array are accepted as a harmless superset."""
    if isinstance(A, (scipy.sparse.csr_matrix, getattr(scipy.sparse, 'csr_array', ()))):
        is_csr = True
    elif isinstance(A, (scipy.sparse.csc_matrix, getattr(scipy.sparse, 'csc_array', ()))):
        is_csr = False
    else:
        raise TypeError("A must be a csr_matrix or a csc_matrix")

    if A.shape[0] != A.shape[1]:
        raise ValueError("A must be a square matrix!")

    A.sort_indices()    # most ILU algorithms require the indices to be ascending
    return A.data, _upcast_indices(A.indices), _upcast_indices(A.indptr), is_csr


def _matrix_from_info(data, indices, indptr, is_csr, rows, cols):
    """reference ilupp/__init__.py:73-82."""
    if is_csr:
        A = scipy.sparse.csr_matrix((data, indices, indptr), shape=(rows, cols), copy=False)
    else:
        A = scipy.sparse.csc_matrix((data, indices, indptr), shape=(rows, cols), copy=False)
    A.has_sorted_indices = True
    return A


def _is_cuda_tensor(x):
    return type(x).__module__.startswith('torch') and getattr(x, 'is_cuda', False)


def _device_args(x, n):
    """(ptr, nrhs, ldx, stream) for a float64 torch CUDA tensor of shape (n,), (n,1)/(n,k) F-ordered
    or (k,n) C-ordered-as-columns; raises the reference's messages otherwise."""
    import torch
    if x.dtype != torch.float64:
        raise RuntimeError("Expected d (d) array for b, got %s!" % (x.dtype,))
    if x.dim() == 1:
        if x.stride(0) != 1:
            raise RuntimeError("Expected contiguous array for b!")
        if x.shape[0] != n:
            raise RuntimeError("vector has wrong size for preconditioner!")
        nrhs, ldx = 1, n
    elif x.dim() == 2:
        if x.shape[0] != n:
            raise RuntimeError("vector has wrong size for preconditioner!")
        if x.stride(0) != 1 and n > 1:
            # also for (n, 1) views such as X[:, :1] of a C-ordered (n, k) tensor: the kernels read n contiguous doubles
            raise RuntimeError("Expected contiguous array for b!" if x.shape[1] == 1 else
                               "Expected column-major (n, k) tensor for multi-RHS apply!")
        nrhs = x.shape[1]
        ldx = x.stride(1) if nrhs > 1 else n
        if nrhs > 1 and ldx < n:
            raise RuntimeError("Expected column-major (n, k) tensor for multi-RHS apply!")
    else:
        raise RuntimeError("Expected 1D array for b!")
    return x.data_ptr(), nrhs, ldx, torch.cuda.current_stream(x.device).cuda_stream


def trim_cache():
    """Return the device memory cached by libilub's allocator (freed blocks kept for reuse) to the driver."""
    _ilupp.trim_cache()


def solve(A, b, rtol=1e-4, atol=1e-4, max_iter=500, threshold=0.1, fill_in=None, params=None, info=False,
          precond="ilu0", device=0):
    """Solve Ax=b with preconditioned BiCGStab (reference ilupp/__init__.py:85-119).

    Same signature and return convention as the reference.  The reference always builds its
    multilevel ILU++ preconditioner here (out of scope); this build runs the same
    BiCGStab loop (SPLIT preconditioning, zero start, tolerances on the preconditioned residual,
    both rtol and atol required) with ``precond`` = "ilu0" (default), entirely on the GPU.  With
    ``precond`` = "ichol0" (symmetric positive definite A) the Krylov method is the reference's preconditioned
    conjugate gradients (cg, iterative_solvers_implementation.h:133-216) with the IChol(0) factor as split
    preconditioner.  ``threshold``/``fill_in``/``params`` are accepted and ignored.
    Raises RuntimeError("did not converge") like the reference.
    """
    Ad, Ai, Ap, Ao = _matrix_fields(A)
    b = np.ascontiguousarray(b, dtype=np.float64)
    if precond == "ilu0":
        sol, it, rtol_out, atol_out = _ilupp.solve(Ad, Ai, Ap, Ao, b, rtol, atol, max_iter, params, device)
    elif precond == "ichol0":
        if b.shape[0] != A.shape[0]:
            raise RuntimeError("right-hand side has wrong size!")
        pr = _ilupp.IChol0Preconditioner(Ad, Ai, Ap, Ao, device)
        sol, it, rtol_out, atol_out = pr.solve(b, rtol, atol, max_iter, True)
    else:
        raise ValueError("precond must be 'ilu0' or 'ichol0'")
    if info:
        return sol, (it, rtol_out, atol_out)
    else:
        return sol


class _BaseWrapper(scipy.sparse.linalg.LinearOperator):
    """reference ilupp/__init__.py:122-168: LinearOperator protocol + apply/apply_trans/total_nnz/factors."""

    def _matvec(self, x):
        y = np.array(x, dtype=np.float64, copy=True).ravel()   # reference: x.copy().ravel()
        self.pr.apply(y)
        return y

    def _rmatvec(self, x):
        y = np.array(x, dtype=np.float64, copy=True).ravel()
        self.pr.apply_trans(y)
        return y

    def _matmat(self, X):
        """Native multi-RHS (the reference inherits scipy's column loop, one apply per column)."""
        Y = np.array(X, dtype=np.float64, order='F', copy=True)
        self.pr.apply_multi(Y, False)
        return Y

    def _rmatmat(self, X):
        Y = np.array(X, dtype=np.float64, order='F', copy=True)
        self.pr.apply_multi(Y, True)
        return Y

    def apply(self, x):
        """Apply the preconditioner to the vector `x` in-place (numpy array, or torch CUDA tensor: zero copy)."""
        if _is_cuda_tensor(x):
            ptr, nrhs, ldx, stream = _device_args(x, self.shape[0])
            self.pr.apply_device(ptr, self.shape[0], nrhs, ldx, stream, False)
        else:
            self.pr.apply(x.ravel())

    def apply_trans(self, x):
        """Apply the transposed preconditioner to the vector `x` in-place."""
        if _is_cuda_tensor(x):
            ptr, nrhs, ldx, stream = _device_args(x, self.shape[0])
            self.pr.apply_device(ptr, self.shape[0], nrhs, ldx, stream, True)
        else:
            self.pr.apply_trans(x.ravel())

    @property
    def total_nnz(self):
        """The total number of nonzeros stored in the factor matrices of the preconditioner."""
        return self.pr.total_nnz

    def factors(self):
        """Return all matrix factors ((L,U) or just (L,)) as a list of sparse matrices."""
        return [_matrix_from_info(*info) for info in self.pr.factors_info()]

    def __repr__(self):
        M, N = self.shape
        if self.dtype is None:
            dt = 'unspecified dtype'
        else:
            dt = 'dtype=' + str(self.dtype)
        return '<%dx%d %s with nnz=%d, %s>' % (M, N, self.__class__.__name__, self.total_nnz, dt)

    # ---- extensions ------------------------------------------------------------------------------
    def refactor(self, A):
        """Numeric re-factorisation with the values of `A` (same pattern as at construction)."""
        if _is_cuda_tensor(A):
            import torch
            self.pr.refactor_device(A.data_ptr(), torch.cuda.current_stream(A.device).cuda_stream)
        else:
            data = A.data if scipy.sparse.issparse(A) else A
            self.pr.refactor(np.ascontiguousarray(data, dtype=np.float64))

    def levels(self, which=0):
        """(level_of_row, nlevels) of solve `which` (0/1: the two solves of apply, 2/3: of apply_trans)."""
        return self.pr.levels(which)

    def level_sets(self, which=0):
        """(level_ptr, level_rows): rows sorted by (level, row)."""
        return self.pr.level_sets(which)

    def solve(self, b, rtol=1e-4, atol=1e-4, max_iter=500, info=False, method=None):
        """Krylov solve for the matrix this preconditioner was built from, with self as split preconditioner:
        method = "bicgstab" or "cg" (default: cg for IChol0Preconditioner, BiCGStab otherwise)."""
        if method is None:
            method = "cg" if isinstance(self, (IChol0Preconditioner, IChol0PreconditionerBatched)) else "bicgstab"
        if method not in ("cg", "bicgstab"):
            raise ValueError("method must be 'cg' or 'bicgstab'")
        sol, it, r, a = self.pr.solve(np.ascontiguousarray(b, dtype=np.float64), rtol, atol, max_iter, method == "cg")
        return (sol, (it, r, a)) if info else sol

    def check_status(self, stream=0):
        """Synchronise `stream` and raise if a kernel of an asynchronous (torch-tensor) call on this object was aborted
        by the watchdog.  (Such a failure is otherwise raised by the next call on the object.)"""
        self.pr.check_status(stream)

    def write_binary(self, filename):
        """Save the factors in the reference's binary matrix format as `<filename>-L.isp` (and `<filename>-R.isp` for
        two-factor preconditioners): the file pair of the reference's ILUC / ILUT `write_binary`
        (preconditioner_implementation.h:960-980, sparse_implementation.h:3839-3860).  Restore with
        `FactorsPreconditioner.read_binary(filename)`."""
        self.pr.write_binary(str(filename))

    def set_option(self, key, value):
        """Set a libilub option for THIS object only (process-wide defaults: ilupp_b200.set_option)."""
        self.pr.set_option(key, int(value))

    def get_option(self, key):
        """The value of a libilub option in force for this object (its own setting, else the process-wide default)."""
        return self.pr.get_option(key)

    @property
    def timings_ms(self):
        """(symbolic, numeric, packing, total) milliseconds of the last factorisation (CUDA events)."""
        return self.pr.timings_ms

    @property
    def stats(self):
        return self.pr.stats


class ILU0Preconditioner(_BaseWrapper):
    """An ILU(0) preconditioner (no fill-in, same sparsity pattern as A).

    Args:
        A: a sparse matrix in CSR or CSC format
    """
    def __init__(self, A, device=0, options=None):
        Ad, Ai, Ap, Ao = _matrix_fields(A)
        _stage_options(options)
        self.pr = _ilupp.ILU0Preconditioner(Ad, Ai, Ap, Ao, device)
        scipy.sparse.linalg.LinearOperator.__init__(self, shape=A.shape, dtype=A.dtype)


class IChol0Preconditioner(_BaseWrapper):
    """An IChol(0) preconditioner (no fill-in, same sparsity pattern as A) for a
    symmetric positive definite matrix.

    Args:
        A: a symmetric sparse matrix in CSR or CSC format
    """
    def __init__(self, A, device=0, options=None):
        Ad, Ai, Ap, Ao = _matrix_fields(A)
        _stage_options(options)
        self.pr = _ilupp.IChol0Preconditioner(Ad, Ai, Ap, Ao, device)
        scipy.sparse.linalg.LinearOperator.__init__(self, shape=A.shape, dtype=A.dtype)


class FactorsPreconditioner(_BaseWrapper):
    """GPU `apply` for triangular factors computed elsewhere: `x = U^-1 L^-1 b` for a lower
    triangular `L` and an upper triangular `U` (one of the two with a unit diagonal), or `x = L^-T L^-1 b` when only `L` (lower triangular, with its
    diagonal) is given.  This is how the factors of the reference's other triangular preconditioners --
    `ICholTPreconditioner`, `ILUTPreconditioner`, `ILUCPreconditioner` (`.factors()`), whose threshold factorisations
    stay on the CPU -- run through the same triangular-solve kernels as ILU(0) / IChol(0).

    Pivoted factors (the reference's `ILUTPPreconditioner` / `ILUCPPreconditioner`, whose `.permutations()` returns
    `(perm_left, perm_right)`; `triangular_solve_perm`, sparse_implementation.h:4166-4258): pass the factors as the
    reference returns them together with `perm_left` (then `L[perm_left, :]` is lower triangular) and / or
    `perm_right` (then `U[:, perm_right]` is upper triangular).

    Args:
        L: sparse lower triangular factor (any scipy format)
        U: sparse upper triangular factor, or None for `L L^T`
        perm_left, perm_right: pivoting permutations (None: identity)
    """
    def __init__(self, L, U=None, device=0, perm_left=None, perm_right=None):
        L = scipy.sparse.csr_matrix(L)
        if L.shape[0] != L.shape[1]:
            raise ValueError('Input matrix must be square!')
        for perm in (perm_left, perm_right):
            if perm is not None and (np.asarray(perm).shape != (L.shape[0],) or
                                     not np.array_equal(np.sort(np.asarray(perm)), np.arange(L.shape[0]))):
                raise ValueError('perm_left / perm_right must be permutations of range(n)!')
        if perm_left is not None and U is None:
            raise ValueError('permutations need both factors!')
        if perm_left is not None:
            L = L[np.asarray(perm_left), :].tocsr()      # the truly triangular factor; apply gathers b[perm_left] first
        if perm_right is not None:
            if U is None:
                raise ValueError('permutations need both factors!')
            U = scipy.sparse.csc_matrix(U)[:, np.asarray(perm_right)].tocsr()   # apply scatters x[perm_right[k]] = z[k] last
        if U is None:
            if scipy.sparse.triu(L, 1).nnz:
                raise ValueError('L must be lower triangular!')
            F = L.copy()
        else:
            U = scipy.sparse.csr_matrix(U)
            if U.shape != L.shape:
                raise ValueError('L and U must have the same shape!')
            if scipy.sparse.triu(L, 1).nnz or scipy.sparse.tril(U, -1).nnz:
                raise ValueError('L must be lower and U upper triangular!')
            ones = np.ones(L.shape[0])
            if np.array_equal(L.diagonal(), ones):
                F = (scipy.sparse.tril(L, -1) + U).tocsr()
            elif np.array_equal(U.diagonal(), ones):
                # pivots on L, unit diagonal on U (what the reference's ILU0 returns for CSC input): the CSC arrays of
                # the pair are the CSR arrays of its transpose U^T L^T = (unit lower)(upper)
                F = (L + scipy.sparse.triu(U, 1)).tocsc()
            else:
                raise ValueError('L or U must have a unit diagonal!')
        F.sum_duplicates()
        F.sort_indices()
        if F.nnz >= 2**31 - 64:
            raise ValueError('factors exceed 32-bit indexing')
        F.indices = F.indices.astype(_index_dtype)
        F.indptr = F.indptr.astype(_index_dtype)
        Fd, Fi, Fp, Fo = _matrix_fields(F)
        make = _ilupp.LLTFactorsPreconditioner if U is None else _ilupp.LUFactorsPreconditioner
        self.pr = make(Fd, Fi, Fp, Fo, device)
        if perm_left is not None or perm_right is not None:
            self.pr.set_permutations(None if perm_left is None else np.asarray(perm_left, dtype=_index_dtype),
                                     None if perm_right is None else np.asarray(perm_right, dtype=_index_dtype))
        scipy.sparse.linalg.LinearOperator.__init__(self, shape=L.shape, dtype=np.dtype(np.float64))


    @classmethod
    def read_binary(cls, filename, device=0):
        """Restore a preconditioner from `<filename>-L.isp` [+ `<filename>-R.isp`] (written by `write_binary` here or by
        the reference's ILUC / ILUT `write_binary`): L U when both files exist, L L^T when only the first does."""
        import os
        L = read_isp(str(filename) + "-L.isp")
        U = read_isp(str(filename) + "-R.isp") if os.path.exists(str(filename) + "-R.isp") else None
        return cls(L, U, device=device)


def read_isp(path):
    """A sparse matrix from the reference's `.isp` binary file (matrix_sparse::read_binary,
    sparse_implementation.h:3814-3837) as a scipy CSR / CSC matrix."""
    return _matrix_from_info(*_ilupp.isp_read(str(path)))


def write_isp(path, A):
    """Write the CSR / CSC matrix `A` as the reference's `.isp` binary file (matrix_sparse::write_binary,
    sparse_implementation.h:3839-3860), byte for byte what the reference writes for the same matrix."""
    Ad, Ai, Ap, Ao = _matrix_fields(A)
    _ilupp.isp_write(str(path), Ad, Ai, Ap, Ao, A.shape[0], A.shape[1])


class _BatchedWrapper(_BaseWrapper):
    def _init(self, factory, A, values, device, options=None):
        _, Ai, Ap, Ao = _matrix_fields(A)
        values = np.ascontiguousarray(values, dtype=np.float64)
        if values.ndim != 2 or values.shape[1] != Ai.shape[0]:
            raise RuntimeError("indices and data should have the same size!")
        self.nsys = values.shape[0]
        _stage_options(options)
        self.pr = factory(values, Ai, Ap, Ao, device)
        N = A.shape[0] * self.nsys
        scipy.sparse.linalg.LinearOperator.__init__(self, shape=(N, N), dtype=np.dtype(np.float64))


class ILU0PreconditionerBatched(_BatchedWrapper):
    """ILU(0) of `nsys` independent systems sharing the pattern of `A` (values[nsys, nnz]); behaves like
    ILU0Preconditioner of the block-diagonal matrix (system s owns rows [s*n, (s+1)*n))."""
    def __init__(self, A, values, device=0, options=None):
        self._init(_ilupp.ILU0PreconditionerBatched, A, values, device, options)


class IChol0PreconditionerBatched(_BatchedWrapper):
    def __init__(self, A, values, device=0, options=None):
        self._init(_ilupp.IChol0PreconditionerBatched, A, values, device, options)


########################################

def ichol0(A, device=0):
    """Compute the L factor of an incomplete Cholesky decomposition without fill-in for the symmetric matrix A."""
    return _matrix_from_info(*_ilupp.ichol0(*_matrix_fields(A), device))


def ilu0(A, device=0):
    """Compute the (L,U) factors of an incomplete LU decomposition without fill-in."""
    return tuple(_matrix_from_info(*mtx) for mtx in _ilupp.ilu0(*_matrix_fields(A), device))
