// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Thin extern "C" glue over the UNMODIFIED reference templates, compiled from
// the sources where they lie under /root/reference/src (see oracle/Makefile).
// It exists so that (a) the C restatement in oracle/ilub_oracle.c can be pinned
// against the real reference on arbitrary inputs, (b) BASELINE config 3
// (ILU0 + BiCGStab, which the reference's Python `solve` cannot reach because it
// hard-wires the multilevel preconditioner, solving_routines_implementation.h:201-203)
// has a real-reference oracle, and (c) bench.py's `--impl reference` arm can time
// the reference's own C++ without Python overhead.
//
// Everything numerical here is the reference's code: ILU0 (ILU0.hpp:69-106),
// IChol0 (IChol.hpp:62-73), indirect_split_triangular_preconditioner /
// indirect_split_triangular_symmetric_preconditioner (preconditioner.h:179-250),
// bicgstab (iterative_solvers_implementation.h:384-529).
#include "ilupp/iluplusplus_interface.cpp"

#include <cstdint>
#include <cstring>
#include <memory>

using namespace iluplusplus;

typedef indirect_split_triangular_preconditioner<Real, matrix, vector> LUPre;
typedef indirect_split_triangular_symmetric_preconditioner<Real, matrix, vector> LLTPre;

namespace {
struct RefHandle {
    int kind;  // 0 = LU (ILU0), 1 = LLT (IChol0)
    std::unique_ptr<LUPre> lu;
    std::unique_ptr<LLTPre> llt;
    Integer n;
};

matrix view(const double* data, const int32_t* indices, const int32_t* indptr, int32_t n, int is_csr) {
    // same non-owning view binding.cpp:85-89 builds
    return matrix(const_cast<double*>(data), const_cast<Integer*>(indices), const_cast<Integer*>(indptr), n, n,
                  is_csr ? ROW : COLUMN, true);
}

const matrix& factor(const RefHandle* h, int which) {
    if (h->kind == 0) return which == 0 ? h->lu->left_matrix() : h->lu->right_matrix();
    return h->llt->left_matrix();
}
}  // namespace

extern "C" {

int ref_index_size() { return (int)sizeof(Integer); }

void* ref_create_ilu0(const double* data, const int32_t* indices, const int32_t* indptr, int32_t n, int is_csr) {
    matrix A = view(data, indices, indptr, n, is_csr);
    matrix L, U;
    ILU0(A, L, U);  // binding.cpp:366-375
    RefHandle* h = new RefHandle();
    h->kind = 0;
    h->n = n;
    h->lu.reset(new LUPre(std::move(L), LOWER_TRIANGULAR, std::move(U), UPPER_TRIANGULAR));
    return h;
}

void* ref_create_ichol0(const double* data, const int32_t* indices, const int32_t* indptr, int32_t n, int is_csr) {
    matrix A = view(data, indices, indptr, n, is_csr);
    matrix L = IChol0(A);  // binding.cpp:377-386
    RefHandle* h = new RefHandle();
    h->kind = 1;
    h->n = n;
    h->llt.reset(new LLTPre(std::move(L), LOWER_TRIANGULAR));
    return h;
}

void ref_destroy(void* p) { delete static_cast<RefHandle*>(p); }

// binding.cpp:237-254 (apply / apply_trans), in place on x
void ref_apply(void* p, int transpose, double* x) {
    RefHandle* h = static_cast<RefHandle*>(p);
    vector y(h->n, x, true);
    if (h->kind == 0)
        h->lu->apply_preconditioner_only(transpose ? TRANSPOSE : ID, y);
    else
        h->llt->apply_preconditioner_only(transpose ? TRANSPOSE : ID, y);
}

int64_t ref_total_nnz(void* p) {
    RefHandle* h = static_cast<RefHandle*>(p);
    return h->kind == 0 ? h->lu->total_nnz() : h->llt->total_nnz();
}

int ref_num_factors(void* p) { return static_cast<RefHandle*>(p)->kind == 0 ? 2 : 1; }

// dims of factor `which` (0 = left/L, 1 = right/U): nnz, is_csr
void ref_factor_dims(void* p, int which, int64_t* nnz, int* is_csr) {
    const matrix& M = factor(static_cast<RefHandle*>(p), which);
    *nnz = M.actual_non_zeroes();
    *is_csr = (M.orientation == ROW);
}

void ref_get_factor(void* p, int which, double* data, int32_t* indices, int32_t* indptr) {
    RefHandle* h = static_cast<RefHandle*>(p);
    const matrix& M = factor(h, which);
    const Integer nnz = M.actual_non_zeroes();
    std::memcpy(data, M.data, sizeof(double) * nnz);
    std::memcpy(indices, M.indices, sizeof(Integer) * nnz);
    std::memcpy(indptr, M.pointer, sizeof(Integer) * (h->n + 1));
}

// factor `which` of the handle dumped by the reference's own matrix_sparse::write_binary
// (sparse_implementation.h:3839-3860): the .isp layout ILUC/ILUT::write_binary use for "<name>-L.isp" / "-R.isp"
// (preconditioner_implementation.h:960-970)
int ref_write_factor_binary(void* p, int which, const char* filename) {
    try {
        factor(static_cast<RefHandle*>(p), which).write_binary(filename);
        return 0;
    } catch (...) {
        return 1;
    }
}
// a matrix given as arrays, dumped / re-read by the reference's write_binary / read_binary (round trip oracle)
int ref_write_matrix_binary(const double* data, const int32_t* indices, const int32_t* indptr, int32_t n, int is_csr,
                            const char* filename) {
    try {
        view(data, indices, indptr, n, is_csr).write_binary(filename);
        return 0;
    } catch (...) {
        return 1;
    }
}
int ref_read_matrix_binary(const char* filename, int32_t* rows, int32_t* cols, int32_t* nnz, int* is_csr, double* data,
                           int32_t* indices, int32_t* indptr, int64_t cap_nnz, int64_t cap_ptr) {
    try {
        matrix M;
        M.read_binary(filename);
        *rows = M.rows();
        *cols = M.columns();
        *nnz = M.actual_non_zeroes();
        *is_csr = (M.orientation == ROW);
        const Integer np = (M.orientation == ROW ? M.rows() : M.columns()) + 1;
        if (*nnz > cap_nnz || np > cap_ptr) return 2;
        std::memcpy(data, M.data, sizeof(double) * *nnz);
        std::memcpy(indices, M.indices, sizeof(Integer) * *nnz);
        std::memcpy(indptr, M.pointer, sizeof(Integer) * np);
        return 0;
    } catch (...) {
        return 1;
    }
}

// y = A x, through the same member bicgstab uses (sparse_implementation.h:3151-3158)
void ref_spmv(const double* data, const int32_t* indices, const int32_t* indptr, int32_t n, int is_csr,
              const double* x, double* y) {
    matrix A = view(data, indices, indptr, n, is_csr);
    vector vx(n, const_cast<double*>(x), true), vy;
    A.matrix_vector_multiplication(ID, vx, vy);
    for (Integer i = 0; i < n; ++i) y[i] = vy[i];
}

// ILU0 (or IChol0 when use_ichol != 0) + bicgstab(SPLIT, zero start, min_iter = 1)
// -- the loop solve_linear_system runs for every split preconditioner
// (solving_routines_implementation.h:56,66,81).  Tolerances are plain numbers here
// and are converted to -log10 exactly as binding.cpp:211-212 does.
// returns 1 on success (reference criterion, iterative_solvers_implementation.h:498)
int ref_bicgstab(void* p, const double* data, const int32_t* indices, const int32_t* indptr, int32_t n, int is_csr,
                 const double* b, double* x, double rtol, double atol, int32_t max_iter, int32_t* iters,
                 double* rel_achieved, double* abs_achieved) {
    RefHandle* h = static_cast<RefHandle*>(p);
    matrix A = view(data, indices, indptr, n, is_csr);
    vector vb(n, const_cast<double*>(b), true);
    vector vx(n, x, true);
    Real rel_tol = -std::log10(rtol), abs_tol = -std::log10(atol);
    Integer it = max_iter;
    bool ok;
    if (h->kind == 0)
        ok = bicgstab<Real, matrix, vector>(*h->lu, SPLIT, A, vb, vx, 1, it, rel_tol, abs_tol, true);
    else
        ok = bicgstab<Real, matrix, vector>(*h->llt, SPLIT, A, vb, vx, 1, it, rel_tol, abs_tol, true);
    *iters = it;
    *rel_achieved = std::pow(10.0, -rel_tol);  // binding.cpp:226
    *abs_achieved = std::pow(10.0, -abs_tol);
    return ok ? 1 : 0;
}

// cg(SPLIT, zero start, min_iter = 1) of the reference (iterative_solvers_implementation.h:133-216) with the IChol0 /
// ILU0 preconditioner of the handle.  The reference's cg returns true whatever the residual; the flag returned here is
// the criterion its bicgstab applies (:498).
int ref_cg(void* p, const double* data, const int32_t* indices, const int32_t* indptr, int32_t n, int is_csr,
           const double* b, double* x, double rtol, double atol, int32_t max_iter, int32_t* iters,
           double* rel_achieved, double* abs_achieved) {
    RefHandle* h = static_cast<RefHandle*>(p);
    matrix A = view(data, indices, indptr, n, is_csr);
    vector vb(n, const_cast<double*>(b), true);
    vector vx(n, x, true);
    Real rel_tol = -std::log10(rtol), abs_tol = -std::log10(atol);
    Integer it = max_iter;
    if (h->kind == 0)
        cg<Real, matrix, vector>(*h->lu, SPLIT, A, vb, vx, 1, it, rel_tol, abs_tol, true);
    else
        cg<Real, matrix, vector>(*h->llt, SPLIT, A, vb, vx, 1, it, rel_tol, abs_tol, true);
    *iters = it;
    *rel_achieved = std::pow(10.0, -rel_tol);
    *abs_achieved = std::pow(10.0, -abs_tol);
    return (*rel_achieved < rtol && *abs_achieved < atol) ? 1 : 0;
}

}  // extern "C"
