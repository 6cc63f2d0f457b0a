// SpMV and the preconditioned BiCGStab loop, device resident.
//
// Replaces generic_matrix_vector_multiplication_addition (sparse_implementation.h:2733-2760), the
// vector_dense BLAS-1 members (:234-279,414-425,661-677) and bicgstab
// (iterative_solvers_implementation.h:384-529) with SPLIT preconditioning
// (preconditioner_implementation.h:143-147,225-227,250-252).
//
// All scalars (rho, alpha, omega, beta, residual norms, iteration counter, done flag) live in
// device memory; dot products are two-stage (per-block partials, the last block to arrive folds
// them in a fixed order, so results are run-to-run deterministic) and the block that finishes a
// reduction also computes the scalars that depend on it.  The ~29 separate vector sweeps of the
// reference's loop become 5 fused kernels per iteration.  The host only polls the done flag
// every `poll` iterations; once it is set every vector kernel is a no-op.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ilub_common.cuh"
#include "ilub_ptx.cuh"

namespace ilub {
namespace {

inline int blocks_for(int64_t work, int threads) { return (int)((work + threads - 1) / threads); }

// ---- SpMV ---------------------------------------------------------------------------------------
// CSR-read arrays of a CSR input: gather form, G lanes per row.  G == 1 accumulates in stored
// order with separately rounded multiply/add => bit-for-bit the reference loop (:2743-2745).
template <int G, bool EXACT>
__global__ void k_spmv_csr(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                           const double* __restrict__ a, const double* __restrict__ x, double* __restrict__ y) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = t / G;
    const int gl = (int)(t % G);
    double acc = 0.0;
    if (row < n) {
        const int b = ptr[row], e = ptr[row + 1];
        for (int p = b + gl; p < e; p += G) acc = add_mul<EXACT>(acc, a[p], __ldg(x + idx[p]));
    }
    if (G > 1) {
#pragma unroll
        for (int s = G / 2; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s, G);
    }
    if (row < n && gl == 0) y[row] = acc;
}

// CSC input: the stored arrays are the CSR of A^T; y = A x is the reference's scatter loop
// (:2751-2753).  Row i of A is gathered from the transposed strict triangles plus the diagonal in
// ascending source-column order, which is exactly the order in which the scatter loop adds into y[i].
template <bool EXACT>
__global__ void k_spmv_csc(int n, const int* __restrict__ dpos, const int* __restrict__ tu_ptr,
                           const int* __restrict__ tu_idx, const int* __restrict__ tu_pos,
                           const int* __restrict__ tl_ptr, const int* __restrict__ tl_idx,
                           const int* __restrict__ tl_pos, const double* __restrict__ a,
                           const double* __restrict__ x, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    for (int t = tu_ptr[i]; t < tu_ptr[i + 1]; ++t) acc = add_mul<EXACT>(acc, a[tu_pos[t]], __ldg(x + tu_idx[t]));
    acc = add_mul<EXACT>(acc, a[dpos[i]], x[i]);
    for (int t = tl_ptr[i]; t < tl_ptr[i + 1]; ++t) acc = add_mul<EXACT>(acc, a[tl_pos[t]], __ldg(x + tl_idx[t]));
    y[i] = acc;
}

// ---- BiCGStab state -------------------------------------------------------------------------------
struct State {
    double rho, alpha, omega, beta, res, initial_res, rel_tol, abs_tol;
    double d0, d1;  // last pair of dot products
    int iter, max_iter, min_iter, done;
    unsigned counter;  // blocks arrived in the current reduction
    int pad;
};

constexpr int RB = 256;        // reduction block size
constexpr int MAXBLK = 1184;   // 148 SMs * 8

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < RB / 32) t = sh[threadIdx.x];
    if (w == 0) {
#pragma unroll
        for (int s = 4; s > 0; s >>= 1) t += __shfl_xor_sync(0xffffffffu, t, s);
    }
    return t;  // valid in thread 0
}

// Two dot products at once; the last block folds the partials and runs `Fin(state, d0, d1)`.
// mode: 0 -> d0 = a.b, d1 = c.d
template <class Fin>
__device__ __forceinline__ void finish_reduction(State* st, double* partial, double p0, double p1, Fin fin) {
    __shared__ double sh[RB / 32];
    __shared__ bool last;
    const double s0 = block_sum(p0, sh);
    const double s1 = block_sum(p1, sh);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s0;
        partial[2 * blockIdx.x + 1] = s1;
        __threadfence();
        const unsigned prev = atomicAdd(&st->counter, 1u);
        last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double a0 = 0.0, a1 = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += RB) {
            a0 += __ldcg(partial + 2 * b);
            a1 += __ldcg(partial + 2 * b + 1);
        }
        const double t0 = block_sum(a0, sh);
        const double t1 = block_sum(a1, sh);
        if (threadIdx.x == 0) {
            st->counter = 0;
            fin(st, t0, t1);
            __threadfence();
        }
    }
}

// initial: r0 = r; p = r; tmp = p; initial_res = res = ||r||
__global__ void __launch_bounds__(RB) k_bicg_init(int n, const double* __restrict__ r, double* __restrict__ r0,
                                                  double* __restrict__ p, double* __restrict__ tmp,
                                                  double* __restrict__ y, State* st, double* partial) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB) {
        const double v = r[i];
        r0[i] = v;
        p[i] = v;
        tmp[i] = v;
        y[i] = 0.0;
        acc += v * v;
    }
    finish_reduction(st, partial, acc, 0.0, [](State* s, double d0, double) {
        s->initial_res = s->res = sqrt(d0);
        s->iter = 0;
        s->done = 0;  // min_iter >= 1: the loop always runs once (:455)
    });
}

// rho = r.r0 ; alpha = rho / (Ap.r0)        (:466-467)
__global__ void __launch_bounds__(RB) k_bicg_alpha(int n, const double* __restrict__ r, const double* __restrict__ r0,
                                                   const double* __restrict__ Ap, State* st, double* partial) {
    if (*(volatile int*)&st->done) return;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB) {
        const double z = r0[i];
        a0 += r[i] * z;
        a1 += Ap[i] * z;
    }
    finish_reduction(st, partial, a0, a1, [](State* s, double d0, double d1) {
        s->rho = d0;
        s->alpha = d0 / d1;
    });
}

// s = r - alpha*Ap ; tmp = s                (:468)
__global__ void __launch_bounds__(RB) k_bicg_s(int n, const double* __restrict__ r, const double* __restrict__ Ap,
                                               double* __restrict__ s, double* __restrict__ tmp, const State* st) {
    if (*(volatile const int*)&st->done) return;
    const double alpha = st->alpha;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB) {
        const double v = r[i] + (-alpha * Ap[i]);
        s[i] = v;
        tmp[i] = v;
    }
}

// omega = (As.s) / (As.As)                  (:478-482)
__global__ void __launch_bounds__(RB) k_bicg_omega(int n, const double* __restrict__ As, const double* __restrict__ s,
                                                   State* st, double* partial) {
    if (*(volatile int*)&st->done) return;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB) {
        const double t = As[i];
        a0 += t * s[i];
        a1 += t * t;
    }
    finish_reduction(st, partial, a0, a1, [](State* s_, double d0, double d1) { s_->omega = d0 / d1; });
}

// y += alpha*p ; y += omega*s ; r = s - omega*As ; d0 = r.r0 ; d1 = r.r      (:484-487,490)
// then beta = (d0/rho)*(alpha/omega), res = sqrt(d1), iter++, loop condition (:455)
__global__ void __launch_bounds__(RB) k_bicg_update(int n, const double* __restrict__ p, const double* __restrict__ s,
                                                    const double* __restrict__ As, const double* __restrict__ r0,
                                                    double* __restrict__ y, double* __restrict__ r, State* st,
                                                    double* partial) {
    if (*(volatile int*)&st->done) return;
    const double alpha = st->alpha, omega = st->omega;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB) {
        const double sv = s[i];
        double yv = y[i];
        yv += alpha * p[i];
        yv += omega * sv;
        y[i] = yv;
        const double rv = sv + (-omega * As[i]);
        r[i] = rv;
        a0 += rv * r0[i];
        a1 += rv * rv;
    }
    finish_reduction(st, partial, a0, a1, [](State* s_, double d0, double d1) {
        s_->beta = (d0 / s_->rho) * (s_->alpha / s_->omega);
        s_->res = sqrt(d1);
        s_->iter += 1;
        s_->d0 = 1.0;  // "p update pending" marker consumed by k_bicg_p
    });
}

// p = beta*(p - omega*Ap) + r ; tmp = p ; then evaluate the loop condition       (:488-489,455)
__global__ void __launch_bounds__(RB) k_bicg_p(int n, double* __restrict__ p, const double* __restrict__ Ap,
                                               const double* __restrict__ r, double* __restrict__ tmp, State* st) {
    if (*(volatile int*)&st->done) return;
    const double beta = st->beta, omega = st->omega;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB) {
        double pv = p[i];
        pv += -omega * Ap[i];
        pv = beta * pv + r[i];
        p[i] = pv;
        tmp[i] = pv;
    }
}

// evaluated after k_bicg_p so that the vector kernels of one iteration all see the same flag
__global__ void k_bicg_check(State* st) {
    if (st->done) return;
    const bool go = (((st->res / st->initial_res > st->rel_tol) || (st->res > st->abs_tol)) &&
                     (st->iter < st->max_iter)) ||
                    (st->iter < st->min_iter);
    if (!go) st->done = 1;
}

// ---- permutations of pivoted factors (triangular_solve_perm, sparse_implementation.h:4166-4258)
__global__ void k_perm_gather(int n, const int* __restrict__ perm, const double* __restrict__ x, double* __restrict__ y) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) y[k] = x[perm[k]];
}
__global__ void k_perm_scatter(int n, const int* __restrict__ perm, const double* __restrict__ x, double* __restrict__ y) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) y[perm[k]] = x[k];
}

// ---- preconditioned CG (SPLIT), iterative_solvers_implementation.h:133-216: two reductions per iteration
// initial: p = r; tmp = p; y = 0; gamma_old = r.r; initial_res = res = ||r||   (:173-177)
__global__ void __launch_bounds__(RB) k_cg_init(int n, const double* __restrict__ r, double* __restrict__ p,
                                                double* __restrict__ tmp, double* __restrict__ y, State* st, double* partial) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB) {
        const double v = r[i];
        p[i] = v;
        tmp[i] = v;
        y[i] = 0.0;
        acc += v * v;
    }
    finish_reduction(st, partial, acc, 0.0, [](State* s, double d0, double) {
        s->rho = d0;  // gamma_old
        s->initial_res = s->res = sqrt(d0);
        s->iter = 0;
        s->done = 0;
    });
}
// alpha = gamma_old / (Ap.p)                 (:182)
__global__ void __launch_bounds__(RB) k_cg_alpha(int n, const double* __restrict__ Ap, const double* __restrict__ p, State* st,
                                                 double* partial) {
    if (*(volatile int*)&st->done) return;
    double a0 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB) a0 += Ap[i] * p[i];
    finish_reduction(st, partial, a0, 0.0, [](State* s, double d0, double) { s->alpha = s->rho / d0; });
}
// y += alpha p ; r -= alpha Ap ; gamma_new = r.r ; beta = gamma_new / gamma_old ; res = ||r||   (:183-189)
__global__ void __launch_bounds__(RB) k_cg_update(int n, const double* __restrict__ p, const double* __restrict__ Ap,
                                                  double* __restrict__ y, double* __restrict__ r, State* st, double* partial) {
    if (*(volatile int*)&st->done) return;
    const double alpha = st->alpha;
    double a0 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB) {
        y[i] += alpha * p[i];
        const double rv = r[i] + (-alpha) * Ap[i];
        r[i] = rv;
        a0 += rv * rv;
    }
    finish_reduction(st, partial, a0, 0.0, [](State* s, double d0, double) {
        s->beta = d0 / s->rho;
        s->rho = d0;
        s->res = sqrt(d0);
        s->iter += 1;
    });
}
// p = beta p + r ; tmp = p                    (:188)
__global__ void __launch_bounds__(RB) k_cg_p(int n, double* __restrict__ p, const double* __restrict__ r, double* __restrict__ tmp,
                                             const State* st) {
    if (*(volatile const int*)&st->done) return;
    const double beta = st->beta;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB) {
        const double pv = beta * p[i] + r[i];
        p[i] = pv;
        tmp[i] = pv;
    }
}

}  // namespace

void permute_gather(ilub_handle* h, const int* perm, double* x, int64_t ldx, int nrhs, cudaStream_t s) {
    h->permbuf.ensure((size_t)h->n);
    for (int c = 0; c < nrhs; ++c) {
        double* xc = x + (int64_t)c * ldx;
        k_perm_gather<<<blocks_for(h->n, 256), 256, 0, s>>>(h->n, perm, xc, h->permbuf.p);
        ILUB_CUDA(cudaMemcpyAsync(xc, h->permbuf.p, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, s));
        h->launches++;
    }
}
void permute_scatter(ilub_handle* h, const int* perm, double* x, int64_t ldx, int nrhs, cudaStream_t s) {
    h->permbuf.ensure((size_t)h->n);
    for (int c = 0; c < nrhs; ++c) {
        double* xc = x + (int64_t)c * ldx;
        k_perm_scatter<<<blocks_for(h->n, 256), 256, 0, s>>>(h->n, perm, xc, h->permbuf.p);
        ILUB_CUDA(cudaMemcpyAsync(xc, h->permbuf.p, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, s));
        h->launches++;
    }
}

void spmv(ilub_handle* h, const double* x, double* y, cudaStream_t s) {
    const Options& opt = options();
    const int n = h->n;
    const bool exact = opt.exact != 0;
    if (h->is_csr) {
        const double avg = (double)h->nnz / (double)n;
        if (exact)
            k_spmv_csr<1, true><<<blocks_for(n, 128), 128, 0, s>>>(n, h->ptr.p, h->idx.p, h->a.p, x, y);
        else if (avg <= 12)
            k_spmv_csr<4, false><<<blocks_for((int64_t)n * 4, 256), 256, 0, s>>>(n, h->ptr.p, h->idx.p, h->a.p, x, y);
        else
            k_spmv_csr<32, false><<<blocks_for((int64_t)n * 32, 256), 256, 0, s>>>(n, h->ptr.p, h->idx.p, h->a.p, x,
                                                                                     y);
    } else {
        build_transposed(h, true, s);
        build_transposed(h, false, s);
        if (exact)
            k_spmv_csc<true><<<blocks_for(n, 128), 128, 0, s>>>(n, h->dpos.p, h->t_upper.tptr.p, h->t_upper.tidx.p,
                                                                 h->t_upper.tpos.p, h->t_lower.tptr.p,
                                                                 h->t_lower.tidx.p, h->t_lower.tpos.p, h->a.p, x, y);
        else
            k_spmv_csc<false><<<blocks_for(n, 128), 128, 0, s>>>(n, h->dpos.p, h->t_upper.tptr.p, h->t_upper.tidx.p,
                                                                  h->t_upper.tpos.p, h->t_lower.tptr.p,
                                                                  h->t_lower.tidx.p, h->t_lower.tpos.p, h->a.p, x, y);
    }
    h->launches++;
    ILUB_CUDA(cudaGetLastError());
}

// slots of the two solves of a non-transposed apply
static void apply_slots(const ilub_handle* h, int& first, int& second) {
    if (h->kind == KIND_ICHOL0) {
        first = 0;
        second = 3;
    } else if (h->is_csr) {
        first = 0;
        second = 1;
    } else {
        first = 2;
        second = 3;
    }
}

int bicgstab(ilub_handle* h, const double* b, double* x, double rtol, double atol, int max_iter, int* iters,
             double* rel, double* abs_, cudaStream_t s) {
    const int n = h->n;
    int sl_left, sl_right;
    apply_slots(h, sl_left, sl_right);
    // workspace: r, r0, p, s, Ap, As, y, tmp  + partials
    const size_t need = (size_t)8 * n;
    h->work.ensure(need);
    h->scal.ensure(2 * MAXBLK + 64);
    double* r = h->work.p;
    double* r0 = r + n;
    double* p = r0 + n;
    double* sv = p + n;
    double* Ap = sv + n;
    double* As = Ap + n;
    double* y = As + n;
    double* tmp = y + n;
    double* partial = h->scal.p;
    State* st = reinterpret_cast<State*>(h->scal.p + 2 * MAXBLK);
    static_assert(sizeof(State) <= 64 * sizeof(double), "state fits");

    State hs;
    memset(&hs, 0, sizeof(hs));
    // tolerance round trip of binding.cpp:211-212 and iterative_solvers_implementation.h:422-423
    hs.rel_tol = exp(-(-log10(rtol)) * log(10.0));
    hs.abs_tol = exp(-(-log10(atol)) * log(10.0));
    hs.max_iter = max_iter;
    hs.min_iter = 1;
    ILUB_CUDA(cudaMemcpyAsync(st, &hs, sizeof(hs), cudaMemcpyHostToDevice, s));

    const int grid = std::min(MAXBLK, std::max(1, blocks_for(n, RB)));
    // r = L^-1 b     (preconditioned_rhs, SPLIT)
    ILUB_CUDA(cudaMemcpyAsync(r, b, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    run_solve(h, sl_left, r, n, 1, s);
    k_bicg_init<<<grid, RB, 0, s>>>(n, r, r0, p, tmp, y, st, partial);
    h->launches++;

    const int poll = 4;
    State out;
    int launched = 0;
    while (true) {
        for (int k = 0; k < poll; ++k) {
            // Ap = L^-1 A U^-1 p
            run_solve(h, sl_right, tmp, n, 1, s);
            spmv(h, tmp, Ap, s);
            run_solve(h, sl_left, Ap, n, 1, s);
            k_bicg_alpha<<<grid, RB, 0, s>>>(n, r, r0, Ap, st, partial);
            k_bicg_s<<<grid, RB, 0, s>>>(n, r, Ap, sv, tmp, st);
            // As = L^-1 A U^-1 s
            run_solve(h, sl_right, tmp, n, 1, s);
            spmv(h, tmp, As, s);
            run_solve(h, sl_left, As, n, 1, s);
            k_bicg_omega<<<grid, RB, 0, s>>>(n, As, sv, st, partial);
            k_bicg_update<<<grid, RB, 0, s>>>(n, p, sv, As, r0, y, r, st, partial);
            k_bicg_p<<<grid, RB, 0, s>>>(n, p, Ap, r, tmp, st);
            k_bicg_check<<<1, 1, 0, s>>>(st);
            h->launches += 6;
            ++launched;
        }
        ILUB_CUDA(cudaMemcpyAsync(&out, st, sizeof(State), cudaMemcpyDeviceToHost, s));
        ILUB_CUDA(cudaStreamSynchronize(s));
        if (out.done || launched > max_iter + poll) break;
    }
    // x = U^-1 y    (adapt_solution, SPLIT)
    ILUB_CUDA(cudaMemcpyAsync(x, y, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    run_solve(h, sl_right, x, n, 1, s);
    ILUB_CUDA(cudaGetLastError());
    *iters = out.iter;
    const bool ok = (out.res / out.initial_res < hs.rel_tol) && (out.res < hs.abs_tol);  // :498
    *rel = pow(10.0, log10(out.res / out.initial_res));                                  // :499 + binding.cpp:226
    *abs_ = pow(10.0, log10(out.res));
    return ok ? 1 : 0;
}

// Preconditioned conjugate gradients with the handle's factors as SPLIT preconditioner (the natural Krylov method for
// IChol(0): L^-1 A L^-T is symmetric positive definite); replaces cg (iterative_solvers_implementation.h:133-216).
int pcg(ilub_handle* h, const double* b, double* x, double rtol, double atol, int max_iter, int* iters, double* rel,
        double* abs_, cudaStream_t s) {
    const int n = h->n;
    int sl_left, sl_right;
    apply_slots(h, sl_left, sl_right);
    h->work.ensure((size_t)5 * n);  // r, p, Ap, y, tmp
    h->scal.ensure(2 * MAXBLK + 64);
    double* r = h->work.p;
    double* p = r + n;
    double* Ap = p + n;
    double* y = Ap + n;
    double* tmp = y + n;
    double* partial = h->scal.p;
    State* st = reinterpret_cast<State*>(h->scal.p + 2 * MAXBLK);
    State hs;
    memset(&hs, 0, sizeof(hs));
    hs.rel_tol = exp(-(-log10(rtol)) * log(10.0));  // :150-151
    hs.abs_tol = exp(-(-log10(atol)) * log(10.0));
    hs.max_iter = max_iter;
    hs.min_iter = 1;
    ILUB_CUDA(cudaMemcpyAsync(st, &hs, sizeof(hs), cudaMemcpyHostToDevice, s));
    const int grid = std::min(MAXBLK, std::max(1, blocks_for(n, RB)));
    ILUB_CUDA(cudaMemcpyAsync(r, b, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    run_solve(h, sl_left, r, n, 1, s);  // r = L^-1 b
    k_cg_init<<<grid, RB, 0, s>>>(n, r, p, tmp, y, st, partial);
    h->launches++;
    const int poll = 4;
    State out;
    int launched = 0;
    while (true) {
        for (int k = 0; k < poll; ++k) {
            run_solve(h, sl_right, tmp, n, 1, s);  // Ap = L^-1 A U^-1 p
            spmv(h, tmp, Ap, s);
            run_solve(h, sl_left, Ap, n, 1, s);
            k_cg_alpha<<<grid, RB, 0, s>>>(n, Ap, p, st, partial);
            k_cg_update<<<grid, RB, 0, s>>>(n, p, Ap, y, r, st, partial);
            k_cg_p<<<grid, RB, 0, s>>>(n, p, r, tmp, st);
            k_bicg_check<<<1, 1, 0, s>>>(st);
            h->launches += 4;
            ++launched;
        }
        ILUB_CUDA(cudaMemcpyAsync(&out, st, sizeof(State), cudaMemcpyDeviceToHost, s));
        ILUB_CUDA(cudaStreamSynchronize(s));
        if (out.done || launched > max_iter + poll) break;
    }
    ILUB_CUDA(cudaMemcpyAsync(x, y, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    run_solve(h, sl_right, x, n, 1, s);  // x = U^-1 y
    ILUB_CUDA(cudaGetLastError());
    *iters = out.iter;
    const bool ok = (out.res / out.initial_res < hs.rel_tol) && (out.res < hs.abs_tol);
    *rel = pow(10.0, log10(out.res / out.initial_res));
    *abs_ = pow(10.0, log10(out.res));
    return ok ? 1 : 0;
}

}  // namespace ilub
