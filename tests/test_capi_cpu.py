"""CPU: the C-ABI library loads and exports every symbol include/ilub.h declares; argument checking and the
host-side mirror of the reference interface behave like the reference (no compute call needs a GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ilupp_b200", "libilub.so")


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ilub.h")).read()
    return sorted(set(re.findall(r"ILUB_API\s+[\w\s\*]+?\b(ilub_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for must in ("ilub_create_ilu0", "ilub_create_ichol0", "ilub_apply", "ilub_destroy", "ilub_total_nnz",
                 "ilub_get_factor", "ilub_get_levels", "ilub_spmv", "ilub_bicgstab", "ilub_create_ilu0_batched",
                 "ilub_refactor", "ilub_last_error", "ilub_index_size"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "libilub.so not built: run __graft_entry__.build()"
    lib = C.CDLL(LIB)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.ilub_index_size.restype = C.c_int
    assert lib.ilub_index_size() == 4                       # reference default Integer (declarations.h:49-53)
    lib.ilub_version.restype = C.c_int
    assert lib.ilub_version() >= 100


def test_argument_errors_do_not_need_a_gpu():
    lib = C.CDLL(LIB)
    lib.ilub_last_error.restype = C.c_char_p
    h = C.c_void_p()
    # null arrays / n == 0 are rejected before any CUDA call
    rc = lib.ilub_create_ilu0(None, None, None, C.c_int32(4), C.c_int64(4), 1, 0, 0, None, C.byref(h))
    assert rc == 1 and b"null" in lib.ilub_last_error()
    d = (C.c_double * 1)(1.0)
    i = (C.c_int32 * 1)(0)
    p = (C.c_int32 * 2)(0, 1)
    rc = lib.ilub_create_ilu0(d, i, p, C.c_int32(0), C.c_int64(0), 1, 0, 0, None, C.byref(h))
    assert rc == 1 and b"matrix has size 0!" in lib.ilub_last_error()   # binding.cpp:80-81
    assert lib.ilub_set_option(b"no_such_option", C.c_int64(1)) == 1
    v = C.c_int64()
    assert lib.ilub_set_option(b"exact", C.c_int64(1)) == 0 and lib.ilub_get_option(b"exact", C.byref(v)) == 0
    assert v.value == 1
    # every documented option key round-trips (include/ilub.h "options"), defaults restored
    for key, probe in ((b"solve_algo", 4), (b"factor_algo", 3), (b"sys_plane", 2), (b"plane_r", 2), (b"plane_wpc", 3),
                       (b"plane_npc", 5), (b"levels_cluster", 0), (b"sys_pdl", 0), (b"spin_limit_ms", 77)):
        old = C.c_int64()
        assert lib.ilub_get_option(key, C.byref(old)) == 0, key
        assert lib.ilub_set_option(key, C.c_int64(probe)) == 0 and lib.ilub_get_option(key, C.byref(v)) == 0
        assert v.value == probe
        assert lib.ilub_set_option(key, old) == 0
    assert lib.ilub_apply(None, 0, None, C.c_int64(1), C.c_int32(1), 0, None) == 1
    # per-handle options: unknown keys and null handles are refused; a staged option is consumed by the next creation
    # of this thread even when that creation fails, and never touches the process-wide default
    assert lib.ilub_stage_option(b"no_such_option", C.c_int64(1)) == 1
    assert lib.ilub_set_handle_option(None, b"exact", C.c_int64(0)) == 1
    assert lib.ilub_get_handle_option(None, b"exact", C.byref(v)) == 1
    assert lib.ilub_stage_option(b"exact", C.c_int64(0)) == 0
    assert lib.ilub_create_ilu0(None, None, None, C.c_int32(4), C.c_int64(4), 1, 0, 0, None, C.byref(h)) == 1
    assert lib.ilub_get_option(b"exact", C.byref(v)) == 0 and v.value == 1
    # factors handed in from elsewhere: same argument checks, no device needed to be refused
    rc = lib.ilub_create_from_factors(0, None, None, None, C.c_int32(4), C.c_int64(4), 1, 0, 0, None, C.byref(h))
    assert rc == 1 and b"null" in lib.ilub_last_error()
    rc = lib.ilub_create_from_factors(7, d, i, p, C.c_int32(1), C.c_int64(1), 1, 0, 0, None, C.byref(h))
    assert rc == 1 and b"kind" in lib.ilub_last_error()


def test_python_layer_errors_match_reference():
    import ilupp_b200 as ilupp
    A = scipy.sparse.random(6, 6, density=0.5, random_state=0, format="coo")
    with pytest.raises(TypeError, match="A must be a csr_matrix or a csc_matrix"):   # ilupp/__init__.py:60-65
        ilupp.ILU0Preconditioner(A)
    with pytest.raises(ValueError, match="A must be a square matrix!"):               # :67-68
        ilupp.ILU0Preconditioner(scipy.sparse.csr_matrix(np.ones((3, 4))))
    B = scipy.sparse.identity(5, format="csr")
    B.indices = B.indices.astype(np.int64)
    B.indptr = B.indptr.astype(np.int64)
    with pytest.raises(TypeError, match="8 bytes per index"):                         # :47-53
        ilupp._matrix_fields(B)
    # in-place sort of the caller's matrix (:70)
    M = scipy.sparse.csr_matrix((np.array([1.0, 2.0, 3.0]), np.array([1, 0, 1], np.int32), np.array([0, 2, 3], np.int32)),
                                shape=(2, 2))
    assert not np.array_equal(M.indices, [0, 1, 1])
    ilupp._matrix_fields(M)
    assert np.array_equal(M.indices, [0, 1, 1]) and np.array_equal(M.data, [2.0, 1.0, 3.0])


def test_binding_buffer_checks_match_reference_messages():
    import ilupp_b200 as ilupp
    m = ilupp._ilupp
    assert m.index_size() == 4
    d = np.ones(3)
    i = np.array([0, 1, 2], np.int32)
    p = np.array([0, 1, 2, 3], np.int32)
    with pytest.raises(RuntimeError, match=r"Expected d \(d\) array for A_data, got f!"):   # binding.cpp:42-51
        m.ILU0Preconditioner(d.astype(np.float32), i, p, True)
    with pytest.raises(RuntimeError, match="Expected integer type with length 4 for A_indices"):
        m.ILU0Preconditioner(d, i.astype(np.int64), p, True)
    with pytest.raises(RuntimeError, match="matrix has size 0!"):                            # :80-81
        m.ILU0Preconditioner(np.zeros(0), np.zeros(0, np.int32), np.zeros(1, np.int32), True)
    with pytest.raises(RuntimeError, match="indices and data should have the same size!"):   # :82-83
        m.ILU0Preconditioner(d[:2], i, p, True)
    with pytest.raises(RuntimeError, match="Expected 1D array for A_data!"):
        m.ILU0Preconditioner(np.ones((3, 1)), i, p, True)
    with pytest.raises(RuntimeError, match="Expected contiguous array for A_data!"):
        m.ILU0Preconditioner(np.ones(6)[::2], i, p, True)


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product path must fail loudly, never compute on the CPU."""
    import ilupp_b200 as ilupp
    if ilupp._ilupp.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        ilupp.ILU0Preconditioner(scipy.sparse.identity(4, format="csr"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ilupp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "ilub_oracle" not in text and "oracle/" not in text, f


def test_factors_preconditioner_argument_checks():
    """SURVEY.md 8(f)1 entry point: shape / triangularity / unit-diagonal checks happen on the host, and without a device
    the call fails loudly instead of falling back"""
    import ilupp_b200 as ilupp
    n = 6
    L = scipy.sparse.tril(scipy.sparse.random(n, n, density=0.5, random_state=1), -1) + scipy.sparse.identity(n)
    U = scipy.sparse.triu(scipy.sparse.random(n, n, density=0.5, random_state=2), 1) + 3.0 * scipy.sparse.identity(n)
    with pytest.raises(ValueError, match="square"):
        ilupp.FactorsPreconditioner(scipy.sparse.csr_matrix(np.ones((3, 4))))
    with pytest.raises(ValueError, match="same shape"):
        ilupp.FactorsPreconditioner(L, scipy.sparse.identity(n + 1, format="csr"))
    with pytest.raises(ValueError, match="lower and U upper"):
        ilupp.FactorsPreconditioner(U, U)
    with pytest.raises(ValueError, match="lower triangular"):
        ilupp.FactorsPreconditioner(U)
    with pytest.raises(ValueError, match="unit diagonal"):
        ilupp.FactorsPreconditioner(L + scipy.sparse.identity(n), U)
    if ilupp._ilupp.device_count() == 0:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            ilupp.FactorsPreconditioner(L, U)


# ---- persistence: the reference's .isp matrix files -------------------------------------
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_isp_files_written_by_the_reference(fmt, tmp_path):
    """tests/golden/isp_*.isp were written by the unmodified reference's matrix_sparse::write_binary
    (scripts/make_golden.py): reading them gives the reference's factors, writing those back gives the same bytes."""
    import ilupp_b200 as ilupp
    g = np.load(os.path.join(GOLD, "nonsym_tridiag_%s.npz" % fmt))
    for side, key in (("L", "ilu0_L"), ("R", "ilu0_U")):
        path = os.path.join(GOLD, "isp_ilu0_tridiag_%s-%s.isp" % (fmt, side))
        M = ilupp.read_isp(path)
        assert scipy.sparse.isspmatrix_csr(M) == bool(g[key + "_is_csr"])
        assert np.array_equal(M.data, g[key + "_data"]) and np.array_equal(M.indices, g[key + "_indices"])
        assert np.array_equal(M.indptr, g[key + "_indptr"])
        out = tmp_path / "again.isp"
        ilupp.write_isp(out, M)
        assert open(out, "rb").read() == open(path, "rb").read()


def test_isp_c_abi_and_errors(tmp_path):
    lib = C.CDLL(LIB)
    lib.ilub_last_error.restype = C.c_char_p
    path = os.path.join(GOLD, "isp_ilu0_tridiag_csr-L.isp").encode()
    rows, cols, nnz, is_csr = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int()
    assert lib.ilub_isp_dims(path, C.byref(rows), C.byref(cols), C.byref(nnz), C.byref(is_csr)) == 0
    assert (rows.value, cols.value, nnz.value, is_csr.value) == (50, 50, 99, 1)
    data, idx, ptr = np.empty(99), np.empty(99, np.int32), np.empty(51, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.ilub_isp_read(path, vp(data), vp(idx), vp(ptr)) == 0
    assert ptr[0] == 0 and ptr[-1] == 99 and idx.max() == 49
    out = str(tmp_path / "copy.isp").encode()
    assert lib.ilub_isp_write(out, rows, cols, 1, vp(data), vp(idx), vp(ptr), C.c_int64(99)) == 0
    assert open(out, "rb").read() == open(path, "rb").read()
    # not there / truncated / not an .isp file: refused with a message, nothing read
    assert lib.ilub_isp_dims(b"/nonexistent/x.isp", C.byref(rows), C.byref(cols), C.byref(nnz), C.byref(is_csr)) == 1
    bad = tmp_path / "bad.isp"
    bad.write_bytes(open(path, "rb").read()[:100])
    assert lib.ilub_isp_read(str(bad).encode(), vp(data), vp(idx), vp(ptr)) == 1 and b"truncated" in lib.ilub_last_error()
    bad.write_bytes(b"\x01" * 64)
    assert lib.ilub_isp_dims(str(bad).encode(), C.byref(rows), C.byref(cols), C.byref(nnz), C.byref(is_csr)) == 1
    assert lib.ilub_write_binary(None, b"x", None) == 1


def test_isp_round_trip_through_the_compiled_reference(tmp_path):
    """ours -> reference read_binary and reference write_binary -> ours, on a random rectangular-free CSR / CSC matrix
    (needs oracle/_ref: the unmodified reference compiled by oracle/Makefile)."""
    import oracle
    if oracle.ref is None:
        pytest.skip("oracle/_ref not built")
    import ilupp_b200 as ilupp
    rng = np.random.default_rng(7)
    for fmt in ("csr", "csc"):
        A = (scipy.sparse.random(40, 40, density=0.15, random_state=3, format=fmt) + scipy.sparse.eye(40)).asformat(fmt)
        A.sort_indices()
        A.data = rng.standard_normal(A.nnz)
        A.indices = A.indices.astype(np.int32); A.indptr = A.indptr.astype(np.int32)
        ours, theirs = tmp_path / ("ours_%s.isp" % fmt), tmp_path / ("ref_%s.isp" % fmt)
        ilupp.write_isp(ours, A)
        oracle.ref.write_matrix_binary(A, theirs)
        assert open(ours, "rb").read() == open(theirs, "rb").read()
        B = oracle.ref.read_matrix_binary(ours, A.nnz, 41)
        assert type(B) is type(A) and np.array_equal(B.data, A.data) and np.array_equal(B.indices, A.indices)
        Cm = ilupp.read_isp(theirs)
        assert type(Cm) is type(A) and np.array_equal(Cm.data, A.data) and np.array_equal(Cm.indptr, A.indptr)
