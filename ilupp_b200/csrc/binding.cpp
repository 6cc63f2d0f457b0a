// pybind11 module `_ilupp` for the B200 path: the same names, argument meaning and error behaviour
// as the hot-path part of the reference's src/binding.cpp, implemented as C++ host code over the
// C ABI of libilub.so (include/ilub.h).  Nothing here does arithmetic.
//
//   reference (src/binding.cpp)                          here
//   index_size                      :279                 index_size            -> ilub_index_size
//   ILU0Preconditioner(...)         :366-375             same                  -> ilub_create_ilu0
//   IChol0Preconditioner(...)       :377-386             same                  -> ilub_create_ichol0
//   ilu0(...) / ichol0(...)         :421-430 / :399-408  same                  -> create + ilub_get_factor
//   solve(...)                      :200-230             same signature, ILU0 preconditioner + ilub_bicgstab
//   GenericLUPreconditioner /
//   GenericLLTPreconditioner        :233-264,358-364     same members          -> ilub_apply, ilub_total_nnz, ...
// Out of scope: every other factory / class of the reference module.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstdint>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <typeinfo>

#include "ilub.h"

extern "C" int ilub_debug_sys_trace(ilub_handle* h, int which_solve, int32_t* dims, long long* trace);
extern "C" int ilub_debug_chunks(ilub_handle* h, int which_solve, int32_t* nchunks, int32_t* rows, long long* trace);

namespace py = pybind11;

namespace {

typedef int32_t Integer;  // declarations.h:49-53 (this build has no INT64_INDICES variant)
typedef double Real;

// ---- buffer validation: same checks and messages as binding.cpp:34-66 ---------------------------
void check_is_contiguous(const py::buffer_info& I, const std::string& name) {
    if (I.ndim != 1) throw std::runtime_error("Expected 1D array for " + name + "!");
    if (I.strides[0] != I.itemsize) throw std::runtime_error("Expected contiguous array for " + name + "!");
}

void check_is_real_array(const py::buffer_info& I, const std::string& name) {
    check_is_contiguous(I, name);
    if (I.format != py::format_descriptor<Real>::format())
        throw std::runtime_error("Expected " + std::string(typeid(Real).name()) + " (" +
                                 py::format_descriptor<Real>::format() + ") array for " + name + ", got " + I.format +
                                 "!");
}

void check_is_int_array(const py::buffer_info& I, const std::string& name) {
    check_is_contiguous(I, name);
    if (!(I.format.length() == 1 && (I.format[0] == 'i' || I.format[0] == 'l' || I.format[0] == 'q') &&
          (I.itemsize == (py::ssize_t)sizeof(Integer))))
        throw std::runtime_error("Expected integer type with length " + std::to_string(sizeof(Integer)) + " for " +
                                 name + ", got " + I.format + "!");
}

[[noreturn]] void raise_status(int st) {
    const std::string msg = ilub_last_error();
    // every C++ exception of the reference's hot path surfaces as RuntimeError through pybind11
    throw std::runtime_error(msg.empty() ? ("libilub error " + std::to_string(st)) : msg);
}
inline void check(int st) {
    if (st != ILUB_OK) raise_status(st);
}

struct MatrixView {  // what make_matrix (binding.cpp:68-90) produces: borrowed arrays + n
    py::buffer_info data, indices, indptr;
    Integer n;
    int64_t nnz;
};

MatrixView make_matrix(py::buffer A_data, py::buffer A_indices, py::buffer A_indptr) {
    MatrixView M;
    M.data = A_data.request();
    check_is_real_array(M.data, "A_data");
    M.indices = A_indices.request();
    check_is_int_array(M.indices, "A_indices");
    M.indptr = A_indptr.request();
    check_is_int_array(M.indptr, "A_indptr");
    M.nnz = M.indices.shape[0];
    if (M.indptr.shape[0] <= 1) throw std::runtime_error("matrix has size 0!");
    if (M.nnz != M.data.shape[0]) throw std::runtime_error("indices and data should have the same size!");
    M.n = (Integer)(M.indptr.shape[0] - 1);
    return M;
}

// ---- the preconditioner object ---------------------------------------------------------------------
struct Precond {
    ilub_handle* h = nullptr;
    Precond() = default;
    explicit Precond(ilub_handle* hh) : h(hh) {}
    Precond(const Precond&) = delete;
    Precond& operator=(const Precond&) = delete;
    Precond(Precond&& o) noexcept : h(o.h) { o.h = nullptr; }
    ~Precond() {
        if (h) ilub_destroy(h);
    }
    int32_t dim() const {
        int32_t n = 0;
        check(ilub_dim(h, &n));
        return n;
    }
};
struct LUPrecond : Precond {  // GenericLUPreconditioner
    using Precond::Precond;
};
struct LLTPrecond : Precond {  // GenericLLTPreconditioner
    using Precond::Precond;
};

template <class P>
P create(bool ichol, py::buffer A_data, py::buffer A_indices, py::buffer A_indptr, bool is_csr, int device) {
    MatrixView M = make_matrix(A_data, A_indices, A_indptr);
    ilub_handle* h = nullptr;
    int st;
    {
        py::gil_scoped_release release;  // binding.cpp:371,382
        st = ichol ? ilub_create_ichol0((const double*)M.data.ptr, (const Integer*)M.indices.ptr,
                                        (const Integer*)M.indptr.ptr, M.n, M.nnz, is_csr ? 1 : 0, ILUB_HOST, device,
                                        nullptr, &h)
                   : ilub_create_ilu0((const double*)M.data.ptr, (const Integer*)M.indices.ptr,
                                      (const Integer*)M.indptr.ptr, M.n, M.nnz, is_csr ? 1 : 0, ILUB_HOST, device,
                                      nullptr, &h);
    }
    check(st);
    return P(h);
}

// factors computed elsewhere (the reference's ICholT / ILUT / ILUC, binding.cpp:388-397,432-470)
template <class P>
P create_from_factors(int kind, py::buffer F_data, py::buffer F_indices, py::buffer F_indptr, bool is_csr, int device) {
    MatrixView M = make_matrix(F_data, F_indices, F_indptr);
    ilub_handle* h = nullptr;
    int st;
    {
        py::gil_scoped_release release;
        st = ilub_create_from_factors(kind, (const double*)M.data.ptr, (const Integer*)M.indices.ptr,
                                      (const Integer*)M.indptr.ptr, M.n, M.nnz, is_csr ? 1 : 0, ILUB_HOST, device, nullptr, &h);
    }
    check(st);
    return P(h);
}

template <class P>
P create_batched(bool ichol, py::array_t<double, py::array::c_style | py::array::forcecast> values,
                 py::buffer A_indices, py::buffer A_indptr, bool is_csr, int device) {
    if (values.ndim() != 2) throw std::runtime_error("Expected 2D array (nsys, nnz) for values!");
    py::buffer_info ii = A_indices.request(), pi = A_indptr.request();
    check_is_int_array(ii, "A_indices");
    check_is_int_array(pi, "A_indptr");
    if (pi.shape[0] <= 1) throw std::runtime_error("matrix has size 0!");
    if (ii.shape[0] != values.shape(1)) throw std::runtime_error("indices and data should have the same size!");
    const Integer n = (Integer)(pi.shape[0] - 1);
    const int64_t nnz = ii.shape[0];
    const int32_t nsys = (int32_t)values.shape(0);
    ilub_handle* h = nullptr;
    int st;
    {
        py::gil_scoped_release release;
        st = ichol ? ilub_create_ichol0_batched(values.data(), (const Integer*)ii.ptr, (const Integer*)pi.ptr, n, nnz,
                                                nsys, is_csr ? 1 : 0, ILUB_HOST, device, nullptr, &h)
                   : ilub_create_ilu0_batched(values.data(), (const Integer*)ii.ptr, (const Integer*)pi.ptr, n, nnz,
                                              nsys, is_csr ? 1 : 0, ILUB_HOST, device, nullptr, &h);
    }
    check(st);
    return P(h);
}

// (data, indices, indptr, is_csr, rows, cols) -- wrap_matrix, binding.cpp:118-133
py::tuple factor_tuple(const Precond& pr, int which) {
    int64_t nnz = 0;
    int is_csr = 1;
    check(ilub_factor_dims(pr.h, which, &nnz, &is_csr));
    const int32_t n = pr.dim();
    py::array_t<Real> data(nnz);
    py::array_t<Integer> indices(nnz), indptr((py::ssize_t)n + 1);
    int st;
    {
        py::gil_scoped_release release;
        st = ilub_get_factor(pr.h, which, data.mutable_data(), indices.mutable_data(), indptr.mutable_data(),
                             ILUB_HOST, nullptr);
    }
    check(st);
    return py::make_tuple(data, indices, indptr, is_csr != 0, n, n);
}

py::list factors_info(const Precond& pr) {
    int count = 0;
    check(ilub_num_factors(pr.h, &count));
    py::list result;
    for (int w = 0; w < count; ++w) result.append(factor_tuple(pr, w));
    return result;
}

void apply_buffer(const Precond& pr, py::buffer x, int transpose) {
    py::buffer_info info = x.request(true);
    check_is_real_array(info, "b");  // make_vector, binding.cpp:92-98
    if (info.shape[0] != pr.dim()) throw std::runtime_error("vector has wrong size for preconditioner!");
    int st;
    {
        py::gil_scoped_release release;  // safe: libilub serialises calls on one handle (per-handle lock)
        st = ilub_apply(pr.h, transpose, (double*)info.ptr, info.shape[0], 1, ILUB_HOST, nullptr);
    }
    check(st);
}

template <class P>
py::class_<P> wrapPreconditioner(py::module& m, const char* classname) {  // binding.cpp:233-264
    return py::class_<P>(m, classname)
        .def("apply", [](const P& pr, py::buffer x) { apply_buffer(pr, x, 0); })
        .def("apply_trans", [](const P& pr, py::buffer x) { apply_buffer(pr, x, 1); })
        .def_property_readonly("total_nnz",
                               [](const P& pr) {
                                   int64_t v = 0;
                                   check(ilub_total_nnz(pr.h, &v));
                                   return v;
                               })
        .def("factors_info", [](const P& pr) { return factors_info(pr); })
        // the generic LU / LL^T classes of the reference report 0 for all three (preconditioner.h:54-57)
        .def_property_readonly("memory_used_calculations", [](const P&) { return 0.0; })
        .def_property_readonly("memory_allocated_calculations", [](const P&) { return 0.0; })
        .def_property_readonly("memory", [](const P&) { return 0.0; })
        .def_property_readonly("exists", [](const P&) { return true; })
        .def_property_readonly("special_info", [](const P&) { return std::string(); })
        .def("print_info",
             [](const P& pr) {
                 int64_t s[16];
                 check(ilub_get_stats(pr.h, s));
                 py::print("B200 zero-fill preconditioner: n =", s[0], "nnz(A) =", s[1], "total_nnz =", s[2]);
             })
        // ---- extensions (no reference counterpart) --------------------------------------------------
        .def_property_readonly("n", [](const P& pr) { return pr.dim(); })
        // in-place apply on nrhs columns in HOST memory: column c starts at ptr + c*ldx doubles
        .def("apply_multi",
             [](const P& pr, py::array_t<double, py::array::f_style> X, bool transpose) {
                 py::buffer_info info = X.request(true);
                 if (info.ndim != 2) throw std::runtime_error("Expected 2D array for X!");
                 if (info.shape[0] != pr.dim()) throw std::runtime_error("vector has wrong size for preconditioner!");
                 int st;
                 {
                     py::gil_scoped_release release;
                     st = ilub_apply(pr.h, transpose ? 1 : 0, (double*)info.ptr, info.shape[0], (int32_t)info.shape[1],
                                     ILUB_HOST, nullptr);
                 }
                 check(st);
             },
             py::arg("X"), py::arg("transpose") = false)
        // in-place apply on device memory (raw pointer, e.g. torch.Tensor.data_ptr()); asynchronous on `stream`
        .def("apply_device",
             [](const P& pr, uintptr_t ptr, int64_t n, int32_t nrhs, int64_t ldx, uintptr_t stream, bool transpose) {
                 if (n != pr.dim()) throw std::runtime_error("vector has wrong size for preconditioner!");
                 check(ilub_apply(pr.h, transpose ? 1 : 0, reinterpret_cast<double*>(ptr), ldx, nrhs, ILUB_DEVICE,
                                  reinterpret_cast<void*>(stream)));
             },
             py::arg("ptr"), py::arg("n"), py::arg("nrhs") = 1, py::arg("ldx") = 0, py::arg("stream") = 0,
             py::arg("transpose") = false)
        .def("set_permutations",
             [](const P& pr, py::object gather_in, py::object scatter_out) {
                 py::array_t<Integer, py::array::c_style | py::array::forcecast> gi, so;
                 const Integer* g = nullptr;
                 const Integer* s = nullptr;
                 if (!gather_in.is_none()) {
                     gi = gather_in.cast<py::array_t<Integer, py::array::c_style | py::array::forcecast>>();
                     if (gi.ndim() != 1 || gi.shape(0) != pr.dim()) throw std::runtime_error("permutation has wrong size!");
                     g = gi.data();
                 }
                 if (!scatter_out.is_none()) {
                     so = scatter_out.cast<py::array_t<Integer, py::array::c_style | py::array::forcecast>>();
                     if (so.ndim() != 1 || so.shape(0) != pr.dim()) throw std::runtime_error("permutation has wrong size!");
                     s = so.data();
                 }
                 check(ilub_set_permutations(pr.h, g, s, ILUB_HOST, nullptr));
             },
             py::arg("gather_in") = py::none(), py::arg("scatter_out") = py::none())
        .def("write_binary",
             [](const P& pr, const std::string& filename) { check(ilub_write_binary(pr.h, filename.c_str(), nullptr)); })
        .def("set_option",
             [](const P& pr, const std::string& k, int64_t v) { check(ilub_set_handle_option(pr.h, k.c_str(), v)); })
        .def("get_option",
             [](const P& pr, const std::string& k) {
                 int64_t v = 0;
                 check(ilub_get_handle_option(pr.h, k.c_str(), &v));
                 return v;
             })
        .def("check_status",
             [](const P& pr, uintptr_t stream) { check(ilub_check_status(pr.h, reinterpret_cast<void*>(stream))); },
             py::arg("stream") = 0)
        .def("levels",
             [](const P& pr, int which) {
                 py::array_t<Integer> lev(pr.dim());
                 int32_t nlev = 0;
                 check(ilub_get_levels(pr.h, which, &nlev, lev.mutable_data(), ILUB_HOST, nullptr));
                 return py::make_tuple(lev, nlev);
             })
        .def("level_sets",
             [](const P& pr, int which) {
                 int32_t nlev = 0;
                 check(ilub_get_levels(pr.h, which, &nlev, nullptr, ILUB_HOST, nullptr));
                 py::array_t<Integer> lptr((py::ssize_t)nlev + 1), lrows(pr.dim());
                 check(ilub_get_level_sets(pr.h, which, lptr.mutable_data(), lrows.mutable_data(), ILUB_HOST, nullptr));
                 return py::make_tuple(lptr, lrows);
             })
        .def("refactor",
             [](const P& pr, py::buffer A_data) {
                 py::buffer_info info = A_data.request();
                 check_is_real_array(info, "A_data");
                 int64_t s[16];
                 check(ilub_get_stats(pr.h, s));
                 if (info.shape[0] != s[1]) throw std::runtime_error("indices and data should have the same size!");
                 int st;
                 {
                     py::gil_scoped_release release;
                     st = ilub_refactor(pr.h, (const double*)info.ptr, ILUB_HOST, nullptr);
                 }
                 check(st);
             })
        .def("refactor_device",
             [](const P& pr, uintptr_t ptr, uintptr_t stream) {
                 check(ilub_refactor(pr.h, reinterpret_cast<const double*>(ptr), ILUB_DEVICE,
                                     reinterpret_cast<void*>(stream)));
             },
             py::arg("ptr"), py::arg("stream") = 0)
        .def("spmv",
             [](const P& pr, py::array_t<double, py::array::c_style | py::array::forcecast> x) {
                 if (x.ndim() != 1 || x.shape(0) != pr.dim()) throw std::runtime_error("right-hand side has wrong size!");
                 py::array_t<Real> y(pr.dim());
                 check(ilub_spmv(pr.h, x.data(), y.mutable_data(), ILUB_HOST, nullptr));
                 return y;
             })
        .def("spmv_device",
             [](const P& pr, uintptr_t x, uintptr_t y, uintptr_t stream) {
                 check(ilub_spmv(pr.h, reinterpret_cast<const double*>(x), reinterpret_cast<double*>(y), ILUB_DEVICE,
                                 reinterpret_cast<void*>(stream)));
             },
             py::arg("x"), py::arg("y"), py::arg("stream") = 0)
        // BiCGStab with this object as split preconditioner; returns the reference's result tuple
        .def("solve",
             [](const P& pr, py::array_t<double, py::array::c_style | py::array::forcecast> b, double rtol, double atol,
                Integer max_iter, bool cg) {
                 if (b.ndim() != 1 || b.shape(0) != pr.dim()) throw std::runtime_error("right-hand side has wrong size!");
                 py::array_t<Real> x(pr.dim());
                 int32_t iters = 0;
                 double rel = 0, ab = 0;
                 int st;
                 {
                     py::gil_scoped_release release;
                     st = (cg ? ilub_pcg : ilub_bicgstab)(pr.h, b.data(), x.mutable_data(), rtol, atol, max_iter, &iters, &rel,
                                                          &ab, ILUB_HOST, nullptr);
                 }
                 if (st == ILUB_ERR_NOT_CONVERGED) throw std::runtime_error("did not converge");  // binding.cpp:228
                 check(st);
                 return py::make_tuple(x, iters, rel, ab);
             },
             py::arg("b"), py::arg("rtol") = 1e-4, py::arg("atol") = 1e-4, py::arg("max_iter") = 500, py::arg("cg") = false)
        .def("solve_device",
             [](const P& pr, uintptr_t b, uintptr_t x, double rtol, double atol, Integer max_iter, uintptr_t stream, bool cg) {
                 int32_t iters = 0;
                 double rel = 0, ab = 0;
                 int st;
                 {
                     py::gil_scoped_release release;
                     st = (cg ? ilub_pcg : ilub_bicgstab)(pr.h, reinterpret_cast<const double*>(b), reinterpret_cast<double*>(x),
                                                          rtol, atol, max_iter, &iters, &rel, &ab, ILUB_DEVICE,
                                                          reinterpret_cast<void*>(stream));
                 }
                 if (st != ILUB_OK && st != ILUB_ERR_NOT_CONVERGED) raise_status(st);
                 return py::make_tuple(iters, rel, ab, st == ILUB_OK);
             },
             py::arg("b"), py::arg("x"), py::arg("rtol") = 1e-4, py::arg("atol") = 1e-4, py::arg("max_iter") = 500,
             py::arg("stream") = 0, py::arg("cg") = false)
        .def("debug_sys_trace",
             [](const P& pr, int which) {
                 int32_t dims[8];
                 check(ilub_debug_sys_trace(pr.h, which, dims, nullptr));
                 py::array_t<long long> trace((size_t)dims[0] * dims[1]);
                 check(ilub_debug_sys_trace(pr.h, which, dims, trace.mutable_data()));
                 py::list d;
                 for (int k = 0; k < 8; ++k) d.append(dims[k]);
                 return py::make_tuple(d, trace);
             })
        .def("debug_chunks",
             [](const P& pr, int which) {
                 int32_t nch = 0;
                 check(ilub_debug_chunks(pr.h, which, &nch, nullptr, nullptr));
                 py::array_t<Integer> rows((py::ssize_t)nch * 32);
                 py::array_t<long long> trace((py::ssize_t)nch * 4);
                 std::fill(trace.mutable_data(), trace.mutable_data() + (size_t)nch * 4, 0LL);
                 check(ilub_debug_chunks(pr.h, which, &nch, rows.mutable_data(), trace.mutable_data()));
                 return py::make_tuple(rows, trace);
             })
        .def_property_readonly("timings_ms",
                               [](const P& pr) {
                                   double t[4];
                                   check(ilub_get_timings(pr.h, t));
                                   return py::make_tuple(t[0], t[1], t[2], t[3]);
                               })
        .def_property_readonly("stats", [](const P& pr) {
            int64_t s[16];
            check(ilub_get_stats(pr.h, s));
            py::dict d;
            d["n"] = s[0];
            d["nnz"] = s[1];
            d["total_nnz"] = s[2];
            d["levels"] = py::make_tuple(s[3], s[4], s[5], s[6]);
            d["launches"] = s[7];
            d["nparts"] = s[8];
            d["part_mode"] = s[9];
            d["grid"] = py::make_tuple(s[10], s[11]);
            d["nchunks"] = s[12];
            d["nslabs"] = s[13];
            d["warps"] = s[14];
            d["smem"] = s[15];
            return d;
        });
}

// Minimal stand-in for the reference's parameter class so that ilupp.solve(..., params=...) keeps its
// signature; the multilevel ILU++ preconditioner those ~70 knobs configure is out of scope here.
struct PrecondParameter {
    double threshold = 1.0;
    int fill_in = 0;
};

}  // namespace

PYBIND11_MODULE(_ilupp, m) {
    m.doc() = "B200-native ILU(0)/IChol(0) preconditioner path (drop-in for the ilupp `_ilupp` hot path)";

    m.def("index_size", []() { return (size_t)ilub_index_size(); });  // binding.cpp:279
    m.def("device_count", []() {
        int c = 0;
        ilub_device_count(&c);
        return c;
    });
    m.def("set_option", [](const std::string& k, int64_t v) { check(ilub_set_option(k.c_str(), v)); });
    // the reference's .isp matrix files (matrix_sparse::read_binary / write_binary, sparse_implementation.h:3814-3860)
    m.def("isp_read", [](const std::string& path) {
        int32_t rows = 0, cols = 0;
        int64_t nnz = 0;
        int is_csr = 1;
        check(ilub_isp_dims(path.c_str(), &rows, &cols, &nnz, &is_csr));
        py::array_t<double> data(nnz);
        py::array_t<int32_t> indices(nnz), indptr((is_csr ? rows : cols) + 1);
        check(ilub_isp_read(path.c_str(), data.mutable_data(), indices.mutable_data(), indptr.mutable_data()));
        return py::make_tuple(data, indices, indptr, is_csr != 0, rows, cols);
    });
    m.def("isp_write", [](const std::string& path, py::array_t<double, py::array::c_style | py::array::forcecast> data,
                          py::array_t<int32_t, py::array::c_style | py::array::forcecast> indices,
                          py::array_t<int32_t, py::array::c_style | py::array::forcecast> indptr, bool is_csr, int32_t rows,
                          int32_t cols) {
        if (indices.size() != data.size()) throw std::runtime_error("indices and data should have the same size!");
        if (indptr.size() != (is_csr ? rows : cols) + 1) throw std::runtime_error("indptr has the wrong size!");
        check(ilub_isp_write(path.c_str(), rows, cols, is_csr ? 1 : 0, data.data(), indices.data(), indptr.data(), data.size()));
    });
    m.def("stage_option", [](const std::string& k, int64_t v) { check(ilub_stage_option(k.c_str(), v)); });
    m.def("get_option", [](const std::string& k) {
        int64_t v = 0;
        check(ilub_get_option(k.c_str(), &v));
        return v;
    });

    py::class_<PrecondParameter>(m, "iluplusplus_precond_parameter")
        .def(py::init<>())
        .def_readwrite("threshold", &PrecondParameter::threshold)
        .def_readwrite("fill_in", &PrecondParameter::fill_in);

    wrapPreconditioner<LUPrecond>(m, "GenericLUPreconditioner");
    wrapPreconditioner<LLTPrecond>(m, "GenericLLTPreconditioner");

    m.def("ILU0Preconditioner",
          [](py::buffer d, py::buffer i, py::buffer p, bool is_csr, int device) {
              return create<LUPrecond>(false, d, i, p, is_csr, device);
          },
          py::arg("A_data"), py::arg("A_indices"), py::arg("A_indptr"), py::arg("is_csr"), py::arg("device") = 0);
    m.def("IChol0Preconditioner",
          [](py::buffer d, py::buffer i, py::buffer p, bool is_csr, int device) {
              return create<LLTPrecond>(true, d, i, p, is_csr, device);
          },
          py::arg("A_data"), py::arg("A_indices"), py::arg("A_indptr"), py::arg("is_csr"), py::arg("device") = 0);
    m.def("LUFactorsPreconditioner",
          [](py::buffer d, py::buffer i, py::buffer p, bool is_csr, int device) {
              return create_from_factors<LUPrecond>(0, d, i, p, is_csr, device);
          },
          py::arg("F_data"), py::arg("F_indices"), py::arg("F_indptr"), py::arg("is_csr"), py::arg("device") = 0);
    m.def("LLTFactorsPreconditioner",
          [](py::buffer d, py::buffer i, py::buffer p, bool is_csr, int device) {
              return create_from_factors<LLTPrecond>(1, d, i, p, is_csr, device);
          },
          py::arg("F_data"), py::arg("F_indices"), py::arg("F_indptr"), py::arg("is_csr"), py::arg("device") = 0);
    m.def("ILU0PreconditionerBatched",
          [](py::array_t<double, py::array::c_style | py::array::forcecast> v, py::buffer i, py::buffer p, bool is_csr,
             int device) { return create_batched<LUPrecond>(false, v, i, p, is_csr, device); },
          py::arg("values"), py::arg("A_indices"), py::arg("A_indptr"), py::arg("is_csr"), py::arg("device") = 0);
    m.def("IChol0PreconditionerBatched",
          [](py::array_t<double, py::array::c_style | py::array::forcecast> v, py::buffer i, py::buffer p, bool is_csr,
             int device) { return create_batched<LLTPrecond>(true, v, i, p, is_csr, device); },
          py::arg("values"), py::arg("A_indices"), py::arg("A_indptr"), py::arg("is_csr"), py::arg("device") = 0);

    m.def("trim_cache", []() { check(ilub_trim_cache()); });
    m.def("ichol0", [](py::buffer d, py::buffer i, py::buffer p, bool is_csr, int device) {  // binding.cpp:399-408
        LLTPrecond pr = create<LLTPrecond>(true, d, i, p, is_csr, device);
        return factor_tuple(pr, 0);
    }, py::arg("A_data"), py::arg("A_indices"), py::arg("A_indptr"), py::arg("is_csr"), py::arg("device") = 0);
    m.def("ilu0", [](py::buffer d, py::buffer i, py::buffer p, bool is_csr, int device) {  // binding.cpp:421-430
        LUPrecond pr = create<LUPrecond>(false, d, i, p, is_csr, device);
        return py::make_tuple(factor_tuple(pr, 0), factor_tuple(pr, 1));
    }, py::arg("A_data"), py::arg("A_indices"), py::arg("A_indptr"), py::arg("is_csr"), py::arg("device") = 0);

    // binding.cpp:200-230.  `param` is accepted for signature compatibility and ignored: the
    // preconditioner is ILU(0) (SURVEY.md 3.3 "gap the build must own").
    m.def("solve",
          [](py::buffer A_data, py::buffer A_indices, py::buffer A_indptr, bool is_csr, py::buffer rhs, double rtol,
             double atol, Integer max_iter, py::object /*param*/, int device) {
              MatrixView M = make_matrix(A_data, A_indices, A_indptr);
              py::buffer_info b = rhs.request();
              check_is_real_array(b, "b");
              if (M.n != b.shape[0]) throw std::runtime_error("right-hand side has wrong size!");
              LUPrecond pr = create<LUPrecond>(false, A_data, A_indices, A_indptr, is_csr, device);
              py::array_t<Real> x(M.n);
              int32_t iters = 0;
              double rel = 0, ab = 0;
              int st;
              {
                  py::gil_scoped_release release;
                  st = ilub_bicgstab(pr.h, (const double*)b.ptr, x.mutable_data(), rtol, atol, max_iter, &iters, &rel,
                                     &ab, ILUB_HOST, nullptr);
              }
              if (st == ILUB_ERR_NOT_CONVERGED) throw std::runtime_error("did not converge");
              check(st);
              return py::make_tuple(x, iters, rel, ab);
          },
          py::arg("A_data"), py::arg("A_indices"), py::arg("A_indptr"), py::arg("is_csr"), py::arg("rhs"),
          py::arg("rtol"), py::arg("atol"), py::arg("max_iter"), py::arg("param") = py::none(), py::arg("device") = 0);
}
