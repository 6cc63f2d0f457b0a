"""GPU: the same hot path called through the raw C ABI (ctypes, plain pointers and sizes) -- what a cgo/JNI/
ctypes FFI of the reference would bind.  Checked against the CPU oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from ilupp_b200 import problems as P
from helpers import relerr, same_pattern

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    L = C.CDLL(os.path.join(ROOT, "ilupp_b200", "libilub.so"))
    L.ilub_last_error.restype = C.c_char_p
    cnt = C.c_int()
    assert L.ilub_device_count(C.byref(cnt)) == 0 and cnt.value > 0
    assert L.ilub_set_option(b"exact", C.c_int64(1)) == 0
    return L


def ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@pytest.mark.parametrize("ichol", [False, True])
def test_create_apply_getfactor_through_c_abi(lib, ichol):
    A = P.poisson3d(12)
    n, nnz = A.shape[0], A.nnz
    data, idx, iptr = A.data.astype(np.float64), A.indices.astype(np.int32), A.indptr.astype(np.int32)
    h = C.c_void_p()
    create = lib.ilub_create_ichol0 if ichol else lib.ilub_create_ilu0
    rc = create(ptr(data, C.c_double), ptr(idx, C.c_int32), ptr(iptr, C.c_int32), C.c_int32(n), C.c_int64(nnz), 1, 0, 0,
                None, C.byref(h))
    assert rc == 0, lib.ilub_last_error()
    try:
        tot = C.c_int64()
        assert lib.ilub_total_nnz(h, C.byref(tot)) == 0
        nf = C.c_int()
        assert lib.ilub_num_factors(h, C.byref(nf)) == 0 and nf.value == (1 if ichol else 2)
        facs = []
        for w in range(nf.value):
            fn, fc = C.c_int64(), C.c_int()
            assert lib.ilub_factor_dims(h, w, C.byref(fn), C.byref(fc)) == 0
            fd, fi, fp = np.empty(fn.value), np.empty(fn.value, np.int32), np.empty(n + 1, np.int32)
            assert lib.ilub_get_factor(h, w, ptr(fd, C.c_double), ptr(fi, C.c_int32), ptr(fp, C.c_int32), 0, None) == 0
            facs.append((fd, fi, fp, fc.value))
        b = P.rhs(n, 0)
        x = b.copy()
        assert lib.ilub_apply(h, 0, ptr(x, C.c_double), C.c_int64(n), C.c_int32(1), 0, None) == 0, lib.ilub_last_error()
        if ichol:
            Lo = oracle.port.ichol0(A)
            assert tot.value == Lo.nnz and np.array_equal(facs[0][0], Lo.data) and np.array_equal(facs[0][1], Lo.indices)
            assert relerr(x, oracle.port.apply_llt(Lo, b)) <= 1e-12
        else:
            Lo, Uo = oracle.port.ilu0(A)
            assert tot.value == Lo.nnz + Uo.nnz
            assert np.array_equal(facs[0][0], Lo.data) and np.array_equal(facs[1][0], Uo.data)
            assert np.array_equal(facs[0][2], Lo.indptr) and np.array_equal(facs[1][1], Uo.indices)
            assert relerr(x, oracle.port.apply_lu(Lo, Uo, b)) <= 1e-12
            xt = b.copy()
            assert lib.ilub_apply(h, 1, ptr(xt, C.c_double), C.c_int64(n), C.c_int32(1), 0, None) == 0
            assert relerr(xt, oracle.port.apply_lu(Lo, Uo, b, True)) <= 1e-12
        # size check of binding.cpp:241-242
        assert lib.ilub_check_vector(h, C.c_int64(n + 1)) == 4
        assert b"vector has wrong size for preconditioner!" in lib.ilub_last_error()
        # levels
        lev = np.empty(n, np.int32)
        nlev = C.c_int32()
        assert lib.ilub_get_levels(h, 0, C.byref(nlev), ptr(lev, C.c_int32), 0, None) == 0
        lo, nlo = oracle.port.levels(A.indptr, A.indices, lower=True)
        assert nlev.value == nlo and np.array_equal(lev, lo)
        st = (C.c_int64 * 16)()
        assert lib.ilub_get_stats(h, st) == 0 and st[0] == n and st[1] == nnz and st[7] > 0
    finally:
        assert lib.ilub_destroy(h) == 0


def test_structure_error_code(lib):
    d = np.ones(3)
    i = np.array([0, 1, 0], np.int32)
    p = np.array([0, 2, 3], np.int32)
    h = C.c_void_p()
    rc = lib.ilub_create_ilu0(ptr(d, C.c_double), ptr(i, C.c_int32), ptr(p, C.c_int32), C.c_int32(2), C.c_int64(3), 1, 0, 0,
                              None, C.byref(h))
    assert rc == 3 and b"missing diagonal" in lib.ilub_last_error() and not h.value


def test_bicgstab_and_spmv_through_c_abi(lib):
    A = P.convdiff3d(16)
    n, nnz = A.shape[0], A.nnz
    data, idx, iptr = A.data.astype(np.float64), A.indices.astype(np.int32), A.indptr.astype(np.int32)
    h = C.c_void_p()
    assert lib.ilub_create_ilu0(ptr(data, C.c_double), ptr(idx, C.c_int32), ptr(iptr, C.c_int32), C.c_int32(n),
                                C.c_int64(nnz), 1, 0, 0, None, C.byref(h)) == 0
    try:
        b = A @ np.ones(n)
        x = np.zeros(n)
        it, rel, ab = C.c_int32(), C.c_double(), C.c_double()
        lib.ilub_bicgstab.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_double,
                                      C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.c_int, C.c_void_p]
        rc = lib.ilub_bicgstab(h, ptr(b, C.c_double), ptr(x, C.c_double), 1e-8, 1e-8, 500, C.byref(it), C.byref(rel),
                               C.byref(ab), 0, None)
        assert rc == 0, lib.ilub_last_error()
        Lo, Uo = oracle.port.ilu0(A)
        xo, ito, _, _, ok = oracle.port.bicgstab(A, b, (Lo, Uo), 1e-8, 1e-8, 500)
        assert ok and abs(it.value - ito) <= 2 and np.abs(x - 1).max() < 1e-6
        rc = lib.ilub_bicgstab(h, ptr(b, C.c_double), ptr(x, C.c_double), 1e-15, 1e-15, 1, C.byref(it), C.byref(rel),
                               C.byref(ab), 0, None)
        assert rc == 5 and b"did not converge" in lib.ilub_last_error()      # binding.cpp:228
        y = np.empty(n)
        v = P.rhs(n, 1)
        assert lib.ilub_spmv(h, ptr(v, C.c_double), ptr(y, C.c_double), 0, None) == 0
        assert np.array_equal(y, oracle.port.spmv(A, v))
    finally:
        lib.ilub_destroy(h)
