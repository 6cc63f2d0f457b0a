"""TEST INFRASTRUCTURE ONLY: ctypes front-ends for the CPU oracle and the compiled reference.

* ``oracle.port``  -- oracle/_build/libilub_oracle.so, the C restatement in ilub_oracle.c
* ``oracle.ref``   -- oracle/_ref/libref_harness.so, the UNMODIFIED reference templates behind
                      extern "C" glue (None if it was never built on this machine)
* ``oracle.ref_module()`` -- the unmodified reference pybind11 module ``_ilupp``

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  ilupp_b200 never does.
"""
import ctypes as C
import importlib.util
import os
import subprocess
import sysconfig

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(ref=True):
    """(Re)build the oracle library and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"] + (["ref"] if ref else []), check=True)


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


def _csr_arrays(A):
    """(ptr, idx, val, is_csr) of a scipy csr/csc matrix, int32/fp64 contiguous, sorted."""
    if not (sp.isspmatrix_csr(A) or sp.isspmatrix_csc(A)):
        raise TypeError("csr or csc expected")
    if not A.has_sorted_indices:
        A = A.copy()
        A.sort_indices()
    return (np.ascontiguousarray(A.indptr, np.int32), np.ascontiguousarray(A.indices, np.int32),
            np.ascontiguousarray(A.data, np.float64), bool(sp.isspmatrix_csr(A)))


def _mk(ptr, idx, val, is_csr, n):
    cls = sp.csr_matrix if is_csr else sp.csc_matrix
    M = cls((val, idx, ptr), shape=(n, n))
    M.has_sorted_indices = True
    return M


# ------------------------------------------------------------------------------------------------
class _Port:
    """The C restatement (kind='port')."""

    def __init__(self):
        path = os.path.join(_HERE, "_build", "libilub_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        i32, i64, f64 = C.c_int32, C.c_int64, C.c_double
        L.oracle_ilu0_values.argtypes = [i32, _i32p, _i32p, _f64p, _f64p]
        L.oracle_ilu0_values.restype = C.c_int
        L.oracle_ilu0_split_sizes.argtypes = [i32, _i32p, _i32p, C.POINTER(i64), C.POINTER(i64)]
        L.oracle_ilu0_split.argtypes = [i32, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
        L.oracle_tril_nnz.argtypes = [i32, _i32p, _i32p]
        L.oracle_tril_nnz.restype = i64
        L.oracle_tril.argtypes = [i32, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
        L.oracle_ichol0_values.argtypes = [i32, _i32p, _i32p, _f64p, _f64p]
        L.oracle_ichol0_values.restype = C.c_int
        for nm in ("oracle_fwd_gather", "oracle_fwd_scatter", "oracle_bwd_gather", "oracle_bwd_scatter"):
            getattr(L, nm).argtypes = [i32, _i32p, _i32p, _f64p, _f64p]
            getattr(L, nm).restype = None
        L.oracle_apply_lu.argtypes = [i32, C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, _f64p]
        L.oracle_apply_llt.argtypes = [i32, _i32p, _i32p, _f64p, _f64p]
        L.oracle_levels.argtypes = [i32, _i32p, _i32p, C.c_int, _i32p]
        L.oracle_levels.restype = i32
        L.oracle_levels_transposed.argtypes = [i32, _i32p, _i32p, C.c_int, _i32p]
        L.oracle_levels_transposed.restype = i32
        L.oracle_spmv.argtypes = [i32, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
        L.oracle_bicgstab.argtypes = [C.c_int, C.c_int, i32, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, C.c_int, _i32p,
                                      _i32p, _f64p, _f64p, _f64p, f64, f64, i32, C.POINTER(i32), C.POINTER(f64),
                                      C.POINTER(f64)]
        L.oracle_bicgstab.restype = C.c_int
        L.oracle_cg.argtypes = L.oracle_bicgstab.argtypes
        L.oracle_cg.restype = C.c_int

    # -- factorisations, returning scipy matrices laid out exactly like ilupp.ilu0 / ilupp.ichol0
    def ilu0(self, A):
        ptr, idx, val, is_csr = _csr_arrays(A)
        n = A.shape[0]
        lu = np.empty_like(val)
        rc = self.lib.oracle_ilu0_values(n, ptr, idx, val, lu)
        if rc:
            raise RuntimeError("missing diagonal in row %d" % (-rc - 1))
        nl, nu = C.c_int64(), C.c_int64()
        self.lib.oracle_ilu0_split_sizes(n, ptr, idx, C.byref(nl), C.byref(nu))
        Lp, Up = np.empty(n + 1, np.int32), np.empty(n + 1, np.int32)
        Li, Lx = np.empty(nl.value, np.int32), np.empty(nl.value)
        Ui, Ux = np.empty(nu.value, np.int32), np.empty(nu.value)
        self.lib.oracle_ilu0_split(n, ptr, idx, lu, Lp, Li, Lx, Up, Ui, Ux)
        if is_csr:
            return _mk(Lp, Li, Lx, True, n), _mk(Up, Ui, Ux, True, n)
        # ILU0.hpp:100-105: we factored A^T; L.interchange(U) and both are relabelled COLUMN
        return _mk(Up, Ui, Ux, False, n), _mk(Lp, Li, Lx, False, n)

    def ichol0(self, A):
        ptr, idx, val, _ = _csr_arrays(A)
        n = A.shape[0]
        nl = self.lib.oracle_tril_nnz(n, ptr, idx)
        Lp, Li, Tx = np.empty(n + 1, np.int32), np.empty(nl, np.int32), np.empty(nl)
        self.lib.oracle_tril(n, ptr, idx, val, Lp, Li, Tx)
        Lx = np.empty(nl)
        rc = self.lib.oracle_ichol0_values(n, Lp, Li, Tx, Lx)
        if rc:
            raise RuntimeError("Matrix passed in the wrong format - should be triangular")
        return _mk(Lp, Li, Lx, True, n)

    # -- apply
    def apply_lu(self, L, U, x, transpose=False):
        Lp, Li, Lx, lcsr = _csr_arrays(L)
        Up, Ui, Ux, _ = _csr_arrays(U)
        y = np.array(x, dtype=np.float64, copy=True).ravel()
        self.lib.oracle_apply_lu(L.shape[0], int(lcsr), int(bool(transpose)), Lp, Li, Lx, Up, Ui, Ux, y)
        return y

    def apply_llt(self, L, x):
        Lp, Li, Lx, _ = _csr_arrays(L)
        y = np.array(x, dtype=np.float64, copy=True).ravel()
        self.lib.oracle_apply_llt(L.shape[0], Lp, Li, Lx, y)
        return y

    def trsv(self, which, M, x):
        ptr, idx, val, _ = _csr_arrays(M)
        y = np.array(x, dtype=np.float64, copy=True).ravel()
        getattr(self.lib, "oracle_" + which)(M.shape[0], ptr, idx, val, y)
        return y

    def levels(self, ptr, idx, lower=True, transposed=False):
        ptr = np.ascontiguousarray(ptr, np.int32)
        idx = np.ascontiguousarray(idx, np.int32)
        lev = np.empty(ptr.size - 1, np.int32)
        f = self.lib.oracle_levels_transposed if transposed else self.lib.oracle_levels
        nlev = f(ptr.size - 1, ptr, idx, int(lower), lev)
        return lev, int(nlev)

    def spmv(self, A, x):
        ptr, idx, val, is_csr = _csr_arrays(A)
        y = np.empty(A.shape[0])
        self.lib.oracle_spmv(A.shape[0], int(is_csr), ptr, idx, val, np.ascontiguousarray(x, np.float64), y)
        return y

    def cg(self, A, b, factors, rtol=1e-8, atol=1e-8, max_iter=2000):
        """reference cg (SPLIT preconditioning); same arguments and result as bicgstab"""
        return self.bicgstab(A, b, factors, rtol, atol, max_iter, fn="oracle_cg")

    def bicgstab(self, A, b, factors, rtol=1e-8, atol=1e-8, max_iter=2000, fn="oracle_bicgstab"):
        """factors = (L, U) from ilu0 or (L,) from ichol0.  Returns (x, iters, rel, abs, ok)."""
        ptr, idx, val, a_csr = _csr_arrays(A)
        n = A.shape[0]
        Lp, Li, Lx, lcsr = _csr_arrays(factors[0])
        if len(factors) == 2:
            Up, Ui, Ux, _ = _csr_arrays(factors[1])
            kind = 0
        else:
            Up, Ui, Ux, kind = Lp, Li, Lx, 1
        x = np.empty(n)
        it, rel, ab = C.c_int32(), C.c_double(), C.c_double()
        ok = getattr(self.lib, fn)(kind, int(lcsr), n, Lp, Li, Lx, Up, Ui, Ux, int(a_csr), ptr, idx, val,
                                   np.ascontiguousarray(b, np.float64), x, rtol, atol, max_iter, C.byref(it),
                                   C.byref(rel), C.byref(ab))
        return x, it.value, rel.value, ab.value, bool(ok)


# ------------------------------------------------------------------------------------------------
class _RefPrecond:
    def __init__(self, ref, handle, n):
        self._ref, self._h, self.n = ref, handle, n

    def __del__(self):
        if self._h:
            self._ref.lib.ref_destroy(self._h)
            self._h = None

    def write_factor_binary(self, which, path):
        """factor `which` dumped by the reference's matrix_sparse::write_binary"""
        f = self._ref.lib.ref_write_factor_binary
        f.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        if f(self._h, which, str(path).encode()) != 0:
            raise RuntimeError("reference write_binary failed")

    def apply(self, x, transpose=False):
        y = np.array(x, dtype=np.float64, copy=True).ravel()
        self._ref.lib.ref_apply(self._h, int(bool(transpose)), y)
        return y

    def apply_inplace(self, y, transpose=False):
        self._ref.lib.ref_apply(self._h, int(bool(transpose)), y)

    @property
    def total_nnz(self):
        return int(self._ref.lib.ref_total_nnz(self._h))

    def factors(self):
        out = []
        for w in range(self._ref.lib.ref_num_factors(self._h)):
            nnz, is_csr = C.c_int64(), C.c_int()
            self._ref.lib.ref_factor_dims(self._h, w, C.byref(nnz), C.byref(is_csr))
            d, i, p = np.empty(nnz.value), np.empty(nnz.value, np.int32), np.empty(self.n + 1, np.int32)
            self._ref.lib.ref_get_factor(self._h, w, d, i, p)
            out.append(_mk(p, i, d, bool(is_csr.value), self.n))
        return out

    def cg(self, A, b, rtol=1e-8, atol=1e-8, max_iter=2000):
        return self.bicgstab(A, b, rtol, atol, max_iter, fn="ref_cg")

    def bicgstab(self, A, b, rtol=1e-8, atol=1e-8, max_iter=2000, fn="ref_bicgstab"):
        ptr, idx, val, is_csr = _csr_arrays(A)
        x = np.zeros(self.n)
        it, rel, ab = C.c_int32(), C.c_double(), C.c_double()
        ok = getattr(self._ref.lib, fn)(self._h, val, idx, ptr, self.n, int(is_csr), np.ascontiguousarray(b, np.float64),
                                        x, rtol, atol, max_iter, C.byref(it), C.byref(rel), C.byref(ab))
        return x, it.value, rel.value, ab.value, bool(ok)


class _Ref:
    """The unmodified reference templates (kind='reference')."""

    def __init__(self, lib):
        L = self.lib = lib
        i32, f64 = C.c_int32, C.c_double
        for nm in ("ref_create_ilu0", "ref_create_ichol0"):
            getattr(L, nm).argtypes = [_f64p, _i32p, _i32p, i32, C.c_int]
            getattr(L, nm).restype = C.c_void_p
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_apply.argtypes = [C.c_void_p, C.c_int, _f64p]
        L.ref_total_nnz.argtypes = [C.c_void_p]
        L.ref_total_nnz.restype = C.c_int64
        L.ref_num_factors.argtypes = [C.c_void_p]
        L.ref_factor_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.ref_get_factor.argtypes = [C.c_void_p, C.c_int, _f64p, _i32p, _i32p]
        L.ref_spmv.argtypes = [_f64p, _i32p, _i32p, i32, C.c_int, _f64p, _f64p]
        L.ref_bicgstab.argtypes = [C.c_void_p, _f64p, _i32p, _i32p, i32, C.c_int, _f64p, _f64p, f64, f64, i32,
                                   C.POINTER(i32), C.POINTER(f64), C.POINTER(f64)]
        L.ref_bicgstab.restype = C.c_int
        if hasattr(L, "ref_cg"):
            L.ref_cg.argtypes = L.ref_bicgstab.argtypes
            L.ref_cg.restype = C.c_int

    def write_matrix_binary(self, A, path):
        """`A` dumped by the reference's own matrix_sparse::write_binary (.isp)"""
        ptr, idx, val, is_csr = _csr_arrays(A)
        f = self.lib.ref_write_matrix_binary
        f.argtypes = [_f64p, _i32p, _i32p, C.c_int32, C.c_int, C.c_char_p]
        if f(val, idx, ptr, A.shape[0], int(is_csr), str(path).encode()) != 0:
            raise RuntimeError("reference write_binary failed")

    def read_matrix_binary(self, path, cap_nnz, cap_ptr):
        """an .isp file read back by the reference's own matrix_sparse::read_binary"""
        f = self.lib.ref_read_matrix_binary
        f.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int), _f64p,
                      _i32p, _i32p, C.c_int64, C.c_int64]
        rows, cols, nnz, is_csr = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int()
        data, idx, ptr = np.empty(cap_nnz), np.empty(cap_nnz, np.int32), np.empty(cap_ptr, np.int32)
        if f(str(path).encode(), C.byref(rows), C.byref(cols), C.byref(nnz), C.byref(is_csr), data, idx, ptr, cap_nnz, cap_ptr) != 0:
            raise RuntimeError("reference read_binary failed")
        np_ = (rows.value if is_csr.value else cols.value) + 1
        return _mk(ptr[:np_].copy(), idx[:nnz.value].copy(), data[:nnz.value].copy(), bool(is_csr.value), rows.value)

    def _create(self, fn, A):
        ptr, idx, val, is_csr = _csr_arrays(A)
        return _RefPrecond(self, fn(val, idx, ptr, A.shape[0], int(is_csr)), A.shape[0])

    def ILU0(self, A):
        return self._create(self.lib.ref_create_ilu0, A)

    def IChol0(self, A):
        return self._create(self.lib.ref_create_ichol0, A)

    def spmv(self, A, x):
        ptr, idx, val, is_csr = _csr_arrays(A)
        y = np.empty(A.shape[0])
        self.lib.ref_spmv(val, idx, ptr, A.shape[0], int(is_csr), np.ascontiguousarray(x, np.float64), y)
        return y


def ref_module():
    """Import the unmodified reference pybind11 extension `_ilupp` built under oracle/_ref."""
    path = os.path.join(_HERE, "_ref", "_ilupp" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(path):
        return None
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec = importlib.util.spec_from_file_location("_ilupp", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    return mod


port = _Port()
_reflib = _load(os.path.join(_HERE, "_ref", "libref_harness.so"))
ref = _Ref(_reflib) if _reflib is not None else None
