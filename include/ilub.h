/*
 * ilub.h -- C ABI of the B200-native ILU(0)/IChol(0) preconditioner path (libilub.so).
 *
 * This is the drop-in boundary: the entry points below are exactly what an FFI for the
 * reference's zero-fill hot path binds.  Each one cites the reference interface it replaces
 * (paths relative to c-f-h/ilupp).  The only in-tree caller is ilupp_b200/csrc/binding.cpp, a
 * pybind11 module exporting the same `_ilupp` names the reference's src/binding.cpp exports.
 *
 * Conventions
 *   - every function returns ILUB_OK (0) or an ilub_status; ilub_last_error() returns a
 *     thread-local, human-readable message for the last failure on the calling thread.
 *     No C++ exception crosses this boundary.
 *   - index arrays are int32 (reference default Integer, declarations.h:49-53), values fp64.
 *   - `where` says where caller-supplied buffers live: ILUB_HOST (pageable or pinned host
 *     memory) or ILUB_DEVICE (device memory on the handle's device).  Input arrays are borrowed
 *     for the duration of the call only (like the non-owning views of binding.cpp:85-89); the
 *     handle owns device copies of everything it needs afterwards.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  With
 *     ILUB_DEVICE buffers calls are asynchronous on that stream; with ILUB_HOST buffers they
 *     return after the result is in host memory.
 *   - a handle is immutable after creation except for its workspace and epoch counters:
 *     serialise apply()/bicgstab() calls per handle (the reference object is const; this one is
 *     not re-entrant).
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     ILUB_ERR_CUDA.
 */
#ifndef ILUB_H
#define ILUB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define ILUB_API __declspec(dllexport)
#else
#define ILUB_API __attribute__((visibility("default")))
#endif

typedef struct ilub_handle ilub_handle;

typedef enum {
    ILUB_OK = 0,
    ILUB_ERR_INVALID = 1,       /* bad argument (null pointer, n <= 0, nrhs <= 0, unknown option ...) */
    ILUB_ERR_CUDA = 2,          /* CUDA runtime error / no device */
    ILUB_ERR_STRUCTURE = 3,     /* unsorted or duplicate column indices, index out of range, or a
                                   structurally missing diagonal (undefined behaviour in the reference,
                                   ILU0.hpp:31,39-44; refused here) */
    ILUB_ERR_SIZE = 4,          /* vector has wrong size for preconditioner! (binding.cpp:241-242) */
    ILUB_ERR_NOT_CONVERGED = 5, /* did not converge (binding.cpp:228) */
    ILUB_ERR_WATCHDOG = 6,      /* a sync-free kernel hit its spin limit (would have dead-locked) */
    ILUB_ERR_ALLOC = 7
} ilub_status;

enum { ILUB_HOST = 0, ILUB_DEVICE = 1 };

/* factor selector for ilub_factor_dims / ilub_get_factor: order of factors_info(), binding.cpp:135-175 */
enum { ILUB_FACTOR_LEFT = 0, ILUB_FACTOR_RIGHT = 1 };

/* which triangular solve's dependency DAG (ilub_get_levels).  Every apply is two solves: */
enum {
    ILUB_SOLVE_FIRST = 0,        /* first solve of apply()       (e.g. L of ILU0-CSR, L of IChol0) */
    ILUB_SOLVE_SECOND = 1,       /* second solve of apply()      (e.g. U of ILU0-CSR, L^T of IChol0) */
    ILUB_SOLVE_TRANS_FIRST = 2,  /* first solve of apply_trans() */
    ILUB_SOLVE_TRANS_SECOND = 3  /* second solve of apply_trans() */
};

/* ---- library -------------------------------------------------------------------------------- */
ILUB_API const char* ilub_last_error(void);
ILUB_API int ilub_version(void);
/* replaces _ilupp.index_size (binding.cpp:279): sizeof(Integer) of this build, always 4 */
ILUB_API int ilub_index_size(void);
ILUB_API int ilub_device_count(int* count);

/* ---- factorisation -------------------------------------------------------------------------- */
/* replaces _ilupp.ILU0Preconditioner / _ilupp.ilu0 (binding.cpp:366-375,421-430):
 * ILU0() of ILU0.hpp:69-106 = compute_ilu0 (:26-66) + L/U split, incl. the CSC relabelling
 * (:100-105).  The matrix must be square with sorted, unique column indices and a full diagonal. */
ILUB_API int ilub_create_ilu0(const double* data, const int32_t* indices, const int32_t* indptr, int32_t n,
                              int64_t nnz, int is_csr, int where, int device, void* stream, ilub_handle** out);
/* replaces _ilupp.IChol0Preconditioner / _ilupp.ichol0 (binding.cpp:377-386,399-408):
 * IChol0() of IChol.hpp:62-73 = natural_triangular_part (sparse_implementation.h:2568-2592) +
 * compute_ichol0 (IChol.hpp:32-59).  The factor is always CSR lower, diagonal last. */
ILUB_API int ilub_create_ichol0(const double* data, const int32_t* indices, const int32_t* indptr, int32_t n,
                                int64_t nnz, int is_csr, int where, int device, void* stream, ilub_handle** out);
/* SURVEY.md 8(f)1 -- apply() for triangular factors computed elsewhere: the reference's ICholTPreconditioner
 * (binding.cpp:388-397, IChol.hpp:77-164), ILUT / ILUC (binding.cpp:432-470,
 * indirect_split_triangular_preconditioner, preconditioner_implementation.h:939-1011) keep their CPU
 * factorisations; their factors run through the same triangular-solve kernels as ILU(0) / IChol(0), i.e. this
 * replaces split_preconditioner::apply_preconditioner_only (preconditioner_implementation.h:91-111) and
 * matrix_sparse::triangular_solve (sparse_implementation.h:4040-4093) for them.
 * kind 0 (L U): `data` holds, on ONE square CSR pattern with sorted unique columns and a full diagonal, L strictly
 * below the diagonal (its unit diagonal is implied) and U on and above it; kind 1 (L L^T): L on the lower triangle,
 * diagonal included (entries above it are ignored).  is_csr = 0: the arrays are the CSC arrays of the factor pair
 * (= CSR of its transpose, exactly how ILU0.hpp:100-105 treats CSC input), i.e. the pair has the unit diagonal on
 * U and the pivots on L -- what the reference's ILU0 returns for CSC matrices.  ilub_refactor on such a handle
 * replaces the factor values; ilub_spmv / ilub_bicgstab are refused (the handle holds no matrix). */
ILUB_API int ilub_create_from_factors(int kind, const double* data, const int32_t* indices, const int32_t* indptr,
                                      int32_t n, int64_t nnz, int is_csr, int where, int device, void* stream,
                                      ilub_handle** out);
/* Pivoting permutations of factors computed elsewhere: replaces matrix_sparse::triangular_solve_perm
 * (sparse_implementation.h:4166-4258) as used by the reference's ILUTP / ILUCP classes
 * (indirect_split_pseudo_triangular_preconditioner, preconditioner_implementation.h:655-757).  With the factors handed
 * to ilub_create_from_factors made truly triangular by the caller (L[perm_left, :] resp. U[:, perm_right]),
 * apply computes  x <- scatter_out(U^-1 L^-1 gather_in(x)),  gather_in(x)[k] = x[gather_in[k]] (the reference's left
 * permutation), scatter_out(z)[scatter_out[k]] = z[k] (its right permutation); apply_trans the transposed operator.
 * Either array may be NULL (identity). */
ILUB_API int ilub_set_permutations(ilub_handle* h, const int32_t* gather_in, const int32_t* scatter_out, int where,
                                   void* stream);
/* numeric re-factorisation with new values on the unchanged pattern; reuses the
 * symbolic analysis (levels, partition, packed layout). */
ILUB_API int ilub_refactor(ilub_handle* h, const double* data, int where, void* stream);
ILUB_API int ilub_destroy(ilub_handle* h);

/* ---- apply ---------------------------------------------------------------------------------- */
/* replaces GenericLUPreconditioner/GenericLLTPreconditioner.apply / .apply_trans
 * (binding.cpp:237-254) = split_preconditioner::apply_preconditioner_only
 * (preconditioner_implementation.h:102-111) over matrix_sparse::triangular_solve
 * (sparse_implementation.h:4040-4087).  In place on x.  nrhs right-hand sides, column c at
 * x + c*ldx (ldx >= n); the reference has no multi-RHS entry (scipy loops columns), nrhs=1 is its
 * exact equivalent. */
ILUB_API int ilub_apply(ilub_handle* h, int transpose, double* x, int64_t ldx, int32_t nrhs, int where,
                        void* stream);
/* Threads: every entry point that takes a handle serialises on a per-handle lock, so one preconditioner may be
 * shared between threads like the reference's (its apply holds the GIL, binding.cpp:237-245).
 * Errors of asynchronous calls: with where = ILUB_DEVICE the kernels run asynchronously on `stream`.  Their
 * watchdog (spin_limit_ms: the persistent sync-free / systolic kernels need all their CTAs co-resident and give up
 * instead of hanging when the GPU is shared and they are not) cannot be reported by the call that launched them:
 * the NEXT call on the handle returns ILUB_ERR_WATCHDOG, or ask explicitly -- ilub_check_status synchronises
 * `stream` and returns ILUB_OK / ILUB_ERR_WATCHDOG (and clears the condition). */
ILUB_API int ilub_check_status(ilub_handle* h, void* stream);
/* give the device memory cached by the library's allocator (freed blocks kept for reuse) back to the driver */
ILUB_API int ilub_trim_cache(void);
/* size check of binding.cpp:241-242 for FFI callers that hold a length: ILUB_ERR_SIZE if len != n */
ILUB_API int ilub_check_vector(const ilub_handle* h, int64_t len);

/* ---- inspection ----------------------------------------------------------------------------- */
/* replaces .total_nnz (binding.cpp:255; preconditioner.h:164,208-209,246-247) */
ILUB_API int ilub_total_nnz(const ilub_handle* h, int64_t* out);
ILUB_API int ilub_dim(const ilub_handle* h, int32_t* n);
/* 2 for ILU0 (L, U), 1 for IChol0 (L): length of factors_info() (binding.cpp:135-175) */
ILUB_API int ilub_num_factors(const ilub_handle* h, int* count);
ILUB_API int ilub_factor_dims(const ilub_handle* h, int which, int64_t* nnz, int* is_csr);
/* replaces .factors_info() / wrap_matrix (binding.cpp:118-133,256): reference-layout arrays of
 * factor `which` into caller-allocated buffers (data[nnz], indices[nnz], indptr[n+1]) */
ILUB_API int ilub_get_factor(ilub_handle* h, int which, double* data, int32_t* indices, int32_t* indptr,
                             int where, void* stream);
/* ---- persistence ------------------------------------------------------------
 * The reference's raw binary dump of a sparse matrix, `.isp` (matrix_sparse::write_binary / read_binary,
 * sparse_implementation.h:3814-3860: int32 rows, cols, pointer_size, nnz, orientation (ROW = 0, COLUMN = 1), then
 * pointer[], indices[] as int32 and data[] as double), byte for byte.
 *   ilub_write_binary   the factors of the handle in reference layout (what ilub_get_factor returns) as
 *                       "<filename>-L.isp" and, for two-factor handles, "<filename>-R.isp" -- the file pair of the
 *                       reference's ILUCPreconditioner / ILUTPreconditioner::write_binary
 *                       (preconditioner_implementation.h:960-980); the reference's generic LU / LL^T classes, which
 *                       ILU0 / IChol0 use, only inherit the stub that throws (preconditioner.h:118-119).
 *   ilub_isp_dims / ilub_isp_read / ilub_isp_write   plain file access for FFI callers (no device needed): read a
 *                       file into caller-allocated arrays (indptr[rows + 1] for a row-oriented matrix, [cols + 1]
 *                       otherwise) and hand them to ilub_create_from_factors to restore a handle. */
ILUB_API int ilub_write_binary(ilub_handle* h, const char* filename, void* stream);
ILUB_API int ilub_isp_dims(const char* path, int32_t* rows, int32_t* cols, int64_t* nnz, int* is_csr);
ILUB_API int ilub_isp_read(const char* path, double* data, int32_t* indices, int32_t* indptr);
ILUB_API int ilub_isp_write(const char* path, int32_t rows, int32_t cols, int is_csr, const double* data,
                            const int32_t* indices, const int32_t* indptr, int64_t nnz);
/* parity hook (no reference equivalent -- its solves are sequential): level of every row in the
 * dependency DAG of one of the triangular solves.  level_of_row[n]; nlevels may be NULL. */
ILUB_API int ilub_get_levels(ilub_handle* h, int which_solve, int32_t* nlevels, int32_t* level_of_row, int where,
                             void* stream);
/* canonical level-set listing: rows sorted by (level, row); level_ptr[nlevels+1], level_rows[n] */
ILUB_API int ilub_get_level_sets(ilub_handle* h, int which_solve, int32_t* level_ptr, int32_t* level_rows,
                                 int where, void* stream);

/* ---- Krylov --------------------------------------------------------------------------------- */
/* y = A x (replaces matrix_sparse::matrix_vector_multiplication,
 * sparse_implementation.h:2733-2760,3151-3158) for the matrix the handle was created from */
ILUB_API int ilub_spmv(ilub_handle* h, const double* x, double* y, int where, void* stream);
/* replaces _ilupp.solve (binding.cpp:200-230) with this handle as the (split) preconditioner:
 * bicgstab(SPLIT, zero start, min_iter 1) of iterative_solvers_implementation.h:384-529.
 * The handle keeps a device copy of A from creation.  On return *iters, *rel_achieved and
 * *abs_achieved are as in the reference's result tuple; ILUB_ERR_NOT_CONVERGED mirrors
 * RuntimeError("did not converge") (x is still written). */
ILUB_API int ilub_bicgstab(ilub_handle* h, const double* b, double* x, double rtol, double atol, int32_t max_iter,
                           int32_t* iters, double* rel_achieved, double* abs_achieved, int where, void* stream);
/* preconditioned conjugate gradients (SPLIT preconditioning with the handle's factors, zero start): replaces cg
 * (iterative_solvers_implementation.h:133-216), the natural Krylov method behind an IChol(0) preconditioner
 * (SURVEY.md 8(f)2).  Same arguments and result convention as ilub_bicgstab; ILUB_ERR_NOT_CONVERGED when the
 * tolerances are not both met (the criterion of iterative_solvers_implementation.h:498). */
ILUB_API int ilub_pcg(ilub_handle* h, const double* b, double* x, double rtol, double atol, int32_t max_iter,
                           int32_t* iters, double* rel_achieved, double* abs_achieved, int where, void* stream);

/* ---- batched independent systems (BASELINE config 5) ---------------------------------------- */
/* nsys systems sharing one sparsity pattern (indices/indptr of ONE system), values[nsys][nnz].
 * The batch is factored as one block-diagonal problem: no dependency ever crosses a system.
 * apply/get_factor/etc. then act on the concatenated system of dimension nsys*n. */
ILUB_API int ilub_create_ilu0_batched(const double* values, const int32_t* indices, const int32_t* indptr,
                                      int32_t n, int64_t nnz, int32_t nsys, int is_csr, int where, int device,
                                      void* stream, ilub_handle** out);
ILUB_API int ilub_create_ichol0_batched(const double* values, const int32_t* indices, const int32_t* indptr,
                                        int32_t n, int64_t nnz, int32_t nsys, int is_csr, int where, int device,
                                        void* stream, ilub_handle** out);

/* ---- options / instrumentation -------------------------------------------------------------- */
/* global options, set before creating handles.  Known keys:
 *   "factor_algo"  0 = auto (default: level loop when levels are few and wide, else sync-free), 1 = one launch per
 *                  level, 2 = sync-free, 3 = level loop (one cooperative launch with grid barriers), 4 = stencil path
 *                  (pivot recurrence on the line-systolic kernel; stencil matrices with "exact" = 0 only, else 3)
 *   "solve_algo"   0 = auto (default: line-systolic / plane-per-warp on structured grids, level loop on deep narrow
 *                  DAGs, cooperative level loop for many moderately wide levels, else chunked sync-free), 1 = one
 *                  launch per level, 2 = chunked sync-free, 3 = level loop (one CTA, one barrier per level),
 *                  4 = cooperative level loop (one launch, grid barrier per level)
 *   "sys_plane"    plane-per-warp kernel for 2-D grids / batches: 0 = off, 1 = auto (default: at least 3 right-hand
 *                  sides or more planes than two waves of CTAs), 2 = whenever applicable; "plane_r", "plane_wpc",
 *                  "plane_npc": right-hand sides per warp (1 | 2), consumer warps per CTA, planes per CTA (0 = auto)
 *   "exact"        1 (default) = IEEE mul/sub/div in the reference's operation order (bit-for-bit
 *                  with a non-FMA build of the reference on lane-per-row chunks); 0 = FMA + reciprocal
 *   "ctas_per_sm", "warps_per_cta", "near_window", "spin_limit_ms" : tuning / watchdog */
ILUB_API int ilub_set_option(const char* key, int64_t value);
ILUB_API int ilub_get_option(const char* key, int64_t* value);
/* Options per handle.  ilub_set_option / ilub_get_option above are the process-wide DEFAULTS; every entry point that
 * works on a handle sees those defaults overlaid with the handle's own settings, so two threads can drive two
 * preconditioners with different settings (the reference objects are const / re-entrant, preconditioner.h:140-260, and
 * carry their parameters themselves) without touching shared state:
 *   ilub_set_handle_option  set `key` for this handle only (takes effect with the next call on the handle; options
 *                           that select a data layout -- "solve_algo", "sys_*", "nparts" ... -- only matter before
 *                           the solve they concern has been analysed, i.e. when staged, below);
 *   ilub_get_handle_option  the value in force for this handle (its own setting, else the default);
 *   ilub_stage_option       thread-local: `key` = `value` for the NEXT handle this thread creates (ilub_create_*),
 *                           which takes the staged settings over as its own before anything is analysed or factorised. */
ILUB_API int ilub_set_handle_option(ilub_handle* h, const char* key, int64_t value);
ILUB_API int ilub_get_handle_option(ilub_handle* h, const char* key, int64_t* value);
ILUB_API int ilub_stage_option(const char* key, int64_t value);
/* milliseconds spent by the last create/refactor on this handle (CUDA events):
 * [0] symbolic analysis, [1] numeric factorisation, [2] solve-layout packing, [3] total */
ILUB_API int ilub_get_timings(const ilub_handle* h, double out_ms[4]);
/* structure statistics: [0] n, [1] nnz(A), [2] total_nnz, [3..6] levels of the four solves (0 if not built),
 * [7] kernels launched by this handle so far; of the first solve of apply(): [8] CTAs (row parts),
 * [9] partition mode (0 contiguous ranges, 2 strips, 3 pencils), [10] x [11] tile grid, [12] 32-row chunks,
 * [13] 32-entry slabs, [14] warps per CTA, [15] dynamic shared memory bytes */
ILUB_API int ilub_get_stats(const ilub_handle* h, int64_t out[16]);

#ifdef __cplusplus
}
#endif
#endif /* ILUB_H */
