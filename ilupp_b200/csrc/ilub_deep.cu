// Level-loop sparse triangular solve for DEEP, NARROW dependency DAGs (BASELINE config 4: irregular matrices whose
// solves have ~10^5 levels of ~10 rows each) -- replaces matrix_sparse::triangular_solve
// (sparse_implementation.h:4040-4087) where neither the line-systolic kernel (structured grids only) nor the chunked
// sync-free kernel (built for wide levels: ~5 us per level on such a DAG) is the right tool.
//
// With ~10 rows per level there is no parallelism to spread over the chip; what matters is the latency of ONE level.
// One CTA walks the levels; a level costs one CTA barrier plus the latency of whatever a row still has to read after
// that barrier, so everything is arranged to leave as little as possible there:
//   * rows are renumbered in solve order (level, row): the right-hand side is permuted into that order by a
//     pre-pass, the solution is kept in solve order (xs) and scattered back by fire-and-forget stores;
//   * the factor is packed in solve order: per position {row, first entry, length}, per entry the dependency's solve
//     POSITION and the coefficient (pre-divided by the diagonal unless `exact`), all read through the read-only
//     path with near-perfect line reuse;
//   * the most recent R solved values live in a shared-memory ring indexed by position, so the dependencies that
//     can be on the critical path (those of the previous level) are a shared-memory read after the barrier;
//   * a group of 8 lanes owns a row; while the CTA is still working on level l, every group prefetches its first row
//     of level l + 1 -- metadata, right-hand side, the first two entries per lane -- and already gathers the
//     dependencies that are final (position below the start of level l), from the ring or from xs through L2.
//     After the barrier only the dependencies that live IN level l remain.
// Arithmetic: `exact` = one lane per row, separately rounded multiply / subtract in the reference's accumulation
// order and a true division (bit-for-bit with the reference); otherwise FMA partial sums over 8 lanes, a shuffle
// reduction and coefficients pre-divided by the diagonal.
#include <algorithm>

#include <cub/cub.cuh>

#include "ilub_common.cuh"
#include "ilub_ptx.cuh"

namespace ilub {
namespace {

constexpr int kDeepThreads = 512;
constexpr int kDeepG = 8;             // lanes per row
constexpr int kDeepRing = 8192;       // solved values kept in shared memory (64 KB)
constexpr int kDeepPf = 2;            // prefetched entries per lane of a group's first row of the next level

inline int blocks_for(int64_t work, int threads) { return (int)((work + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------ analysis
__global__ void k_deep_inverse(int n, const int* __restrict__ rows, int* __restrict__ pos_of_row) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) pos_of_row[rows[p]] = p;
}
__global__ void k_deep_len(int n, SrcDev S, const int* __restrict__ rows, int* __restrict__ len) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) {
        const int r = rows[p];
        len[p] = src_end(S, r) - src_beg(S, r);
    }
}
// per position: meta {row, first entry, length, diagonal source}; per entry: dependency position and value source,
// in the order the reference accumulates them
__global__ void k_deep_fill(int n, SrcDev S, int unit_diag, const int* __restrict__ rows, const int* __restrict__ pos_of_row,
                            const int* __restrict__ eptr, int4* __restrict__ meta, int* __restrict__ dpos,
                            int* __restrict__ dsrc) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int p = (int)(t >> 3), gl = (int)(t & 7);
    if (p >= n) return;
    const int r = rows[p];
    const int b = src_beg(S, r), e = src_end(S, r);
    const int q0 = eptr[p];
    if (gl == 0) meta[p] = make_int4(r, q0, e - b, unit_diag ? -1 : S.diag_pos[r]);
    for (int j = gl; j < e - b; j += 8) {
        const int q = S.reverse ? e - 1 - j : b + j;
        dpos[q0 + j] = pos_of_row[S.dep_idx[q]];
        dsrc[q0 + j] = S.val_pos ? S.val_pos[q] : q;
    }
}
__global__ void k_deep_pack(int n, int recip, const int4* __restrict__ meta, const int* __restrict__ dsrc,
                            const double* __restrict__ lu, double* __restrict__ val, double* __restrict__ diag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int p = (int)(t >> 3), gl = (int)(t & 7);
    if (p >= n) return;
    const int4 m = meta[p];
    const double d = m.w >= 0 ? lu[m.w] : 1.0;
    const double dinv = recip ? 1.0 / d : 1.0;
    if (gl == 0) diag[p] = recip ? dinv : d;
    for (int j = gl; j < m.z; j += 8) val[m.y + j] = recip ? lu[dsrc[m.y + j]] * dinv : lu[dsrc[m.y + j]];
}
__global__ void k_deep_permute_in(int n, const int4* __restrict__ meta, const double* __restrict__ x, double* __restrict__ bs) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) bs[p] = x[meta[p].x];
}

// ------------------------------------------------------------------------------------------------ kernel
struct DeepParams {
    int n, nlevels, nent;
    const int* lvl_ptr;
    const int4* meta;
    const int* dpos;
    const double* val;
    const double* diag;
    const double* bs;   // right-hand side in solve order
    double* xs;         // solution in solve order
    double* x;          // solution in row order (in/out vector of the solve)
    int unit_diag;
};

__device__ __forceinline__ unsigned touch(const void* p) {
    unsigned v;
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// value of the solved position p while positions [.., wend) are being written: the ring is safe for p >= wend - R
__device__ __forceinline__ double deep_x(const DeepParams& P, const double* ring, int p, int wend) {
    if (p >= wend - kDeepRing) return ring[p & (kDeepRing - 1)];
    return __ldcg(P.xs + p);
}

template <bool EXACT>
__global__ void __launch_bounds__(kDeepThreads, 1) k_trsv_deep(const DeepParams P) {
    extern __shared__ double ring[];
    const int lane = threadIdx.x & 31;
    const int gl = threadIdx.x % kDeepG;
    const int g = threadIdx.x / kDeepG;
    const int ng = (kDeepThreads - 32) / kDeepG;  // groups of the compute warps (the last warp only prefetches)
    const unsigned gmask = 0xffu << ((lane / kDeepG) * kDeepG);

    // finish one row whose first kDeepPf * 8 entries may already be (partly) accumulated
    auto finish = [&](int pos, const int4 m, double acc, int jdone, int wend) {
        if (EXACT) {
            if (gl == 0) {
                for (int j = 0; j < m.z; ++j)
                    acc = __dsub_rn(acc, __dmul_rn(__ldg(P.val + m.y + j), deep_x(P, ring, __ldg(P.dpos + m.y + j), wend)));
                if (!P.unit_diag) acc = __ddiv_rn(acc, __ldg(P.diag + pos));
            }
        } else {
            for (int j = jdone + gl; j < m.z; j += kDeepG)
                acc = fma(-__ldg(P.val + m.y + j), deep_x(P, ring, __ldg(P.dpos + m.y + j), wend), acc);
            acc += __shfl_xor_sync(gmask, acc, 4, kDeepG);
            acc += __shfl_xor_sync(gmask, acc, 2, kDeepG);
            acc += __shfl_xor_sync(gmask, acc, 1, kDeepG);
        }
        if (gl == 0) {
            ring[pos & (kDeepRing - 1)] = acc;
            __stcg(P.xs + pos, acc);
            P.x[m.x] = acc;
        }
    };

    // ---- the last warp keeps the read-only streams of the coming levels in L1 by touching their lines ahead of the
    //      level loop (plain loads: they allocate in L1, unlike prefetch hints): a level must never wait for a line
    //      of metadata, entries or right-hand side from L2 / HBM.  It runs free of the per-level barrier.
    __shared__ int prog[3];  // level loop position: end of the current level, its first entry, its number
    if (threadIdx.x < 3) prog[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x >= kDeepThreads - 32) {
        int tpos = 0, tq = 0;
        unsigned sink = 0;
        while (true) {
            const int cur = *(volatile int*)&prog[0], curq = *(volatile int*)&prog[1];
            if (cur >= P.n) break;
            const int wantp = min(P.n, cur + 1024), wantq = min(P.nent, curq + 6144);
            if (tpos < wantp) {  // 256 positions per round: meta 16 B, bs / diag 8 B per position
                sink += touch(P.meta + min(P.n - 1, tpos + lane * 8));
                if (lane < 16) {
                    sink += touch(P.bs + min(P.n - 1, tpos + lane * 16));
                    sink += touch(P.diag + min(P.n - 1, tpos + lane * 16));
                }
                tpos += 256;
            }
            if (tq < wantq) {    // 512 entries per round: dpos 4 B, val 8 B per entry
                if (lane < 16) sink += touch(P.dpos + min(P.nent - 1, tq + lane * 32));
                sink += touch(P.val + min(P.nent - 1, tq + lane * 16));
                tq += 512;
            }
            if (lane == 0) sink += touch(P.lvl_ptr + min(P.nlevels, *(volatile int*)&prog[2] + 64));
            if (tpos >= wantp && tq >= wantq) __nanosleep(200);
        }
        if (sink == 0x9e3779b9u) prog[0] = (int)sink;  // keep the loads alive
        return;
    }
    int lb = __ldg(P.lvl_ptr), le = __ldg(P.lvl_ptr + 1);
    // prefetched state of this group's first row of the level about to be resolved
    int4 pm = make_int4(-1, 0, 0, 0);
    int ppos = -1;
    double pacc = 0.0;
    // per prefetched entry: coefficient, and either the dependency's value (requested before the barrier, consumed
    // after it) or -- if the dependency lives in the level that was being resolved at that time -- its position
    int dp[kDeepPf];
    double dv[kDeepPf], dx[kDeepPf];
#pragma unroll
    for (int k = 0; k < kDeepPf; ++k) dp[k] = -1, dv[k] = 0.0, dx[k] = 0.0;
    if (lb + g < le) {  // level 0 has no dependencies; prefetch trivially
        ppos = lb + g;
        pm = __ldg(P.meta + ppos);
        pacc = gl == 0 ? (EXACT ? __ldg(P.bs + ppos) : __ldg(P.bs + ppos) * __ldg(P.diag + ppos)) : 0.0;
    }

    for (int l = 0; l < P.nlevels; ++l) {
        const int lnb = le;                                                    // next level [lnb, lne)
        const int lne = l + 1 < P.nlevels ? __ldg(P.lvl_ptr + l + 2) : le;
        // ---- this group's first row of level l: only what could not be done before the barrier
        if (ppos >= 0) {
            double acc = pacc;
            if (!EXACT) {
#pragma unroll
                for (int k = 0; k < kDeepPf; ++k) {
                    if (dp[k] >= 0) dx[k] = deep_x(P, ring, dp[k], le);
                    acc = fma(-dv[k], dx[k], acc);
                }
            }
            finish(ppos, pm, acc, EXACT ? 0 : kDeepPf * kDeepG, le);
        }
        // ---- further rows of a wide level
        for (int pos = lb + g + ng; pos < le; pos += ng) {
            const int4 m = __ldg(P.meta + pos);
            const double acc = gl == 0 ? (EXACT ? __ldg(P.bs + pos) : __ldg(P.bs + pos) * __ldg(P.diag + pos)) : 0.0;
            finish(pos, m, acc, 0, le);
        }
        // ---- prefetch the first row of level l + 1 and gather what is already final (positions < lb ... < le are
        //      being written right now: dependencies at positions >= lb wait for the barrier)
        ppos = -1;
#pragma unroll
        for (int k = 0; k < kDeepPf; ++k) dp[k] = -1, dv[k] = 0.0, dx[k] = 0.0;
        if (lnb + g < lne) {
            ppos = lnb + g;
            pm = __ldg(P.meta + ppos);
            pacc = gl == 0 ? (EXACT ? __ldg(P.bs + ppos) : __ldg(P.bs + ppos) * __ldg(P.diag + ppos)) : 0.0;
            if (!EXACT) {
#pragma unroll
                for (int k = 0; k < kDeepPf; ++k) {
                    const int j = gl + k * kDeepG;
                    if (j < pm.z) {
                        const int p = __ldg(P.dpos + pm.y + j);
                        const double v = __ldg(P.val + pm.y + j);
                        dv[k] = v;
                        if (p < lb)
                            dx[k] = deep_x(P, ring, p, le);  // final already: requested now, used after the barrier
                        else
                            dp[k] = p;
                    }
                }
            }
        }
        if (threadIdx.x == 0) {  // where the level loop is, for the toucher warp
            *(volatile int*)&prog[0] = lne;
            *(volatile int*)&prog[2] = l + 1;
            *(volatile int*)&prog[1] = lne < P.n ? __ldg(P.meta + lne).y : P.nent;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kDeepThreads - 32) : "memory");  // the compute warps only
        lb = lnb;
        le = lne;
    }
}

}  // namespace

// ================================================================================================== host
bool deep_wanted(const ilub_handle* h, int slot) {
    const TriSolve& T = h->solve[slot];
    const int64_t algo = options().solve_algo;
    if (algo == 3) return true;
    if (algo != 0 || T.sys.built) return false;
    return (int64_t)T.nlevels * 64 > (int64_t)h->n && T.nlevels > 64;  // fewer than 64 rows per level on average
}

void analyse_deep(ilub_handle* h, int slot, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    DeepSolve& D = T.deep;
    if (D.built) return;
    const int n = h->n;
    const SolveSource src = solve_source(h, slot);
    const SrcDev S = to_dev(src);
    build_level_sets(h, slot, s);
    DevBuf<int> pos_of_row, len;
    pos_of_row.alloc(n);
    len.alloc(n + 1);
    D.eptr.alloc(n + 1);
    k_deep_inverse<<<blocks_for(n, 256), 256, 0, s>>>(n, T.lvl_rows.p, pos_of_row.p);
    ILUB_CUDA(cudaMemsetAsync(len.p + n, 0, sizeof(int), s));
    k_deep_len<<<blocks_for(n, 256), 256, 0, s>>>(n, S, T.lvl_rows.p, len.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, len.p, D.eptr.p, n + 1, s);
    DevBuf<unsigned char> tmp;
    tmp.alloc(tb);
    ILUB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, len.p, D.eptr.p, n + 1, s));
    int nent = 0;
    ILUB_CUDA(cudaMemcpyAsync(&nent, D.eptr.p + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    D.nent = nent;
    D.meta.alloc(n);
    D.dpos.alloc(std::max(1, nent));
    D.dsrc.alloc(std::max(1, nent));
    D.val.alloc(std::max(1, nent));
    D.diag.alloc(n);
    D.bs.alloc(n);
    D.xs.alloc(n);
    k_deep_fill<<<blocks_for((int64_t)n * 8, 256), 256, 0, s>>>(n, S, src.unit_diag ? 1 : 0, T.lvl_rows.p, pos_of_row.p,
                                                               D.eptr.p, D.meta.p, D.dpos.p, D.dsrc.p);
    h->launches += 4;
    ILUB_CUDA(cudaStreamSynchronize(s));  // temporaries die here
    D.packed_exact = -1;
    D.built = true;
}

void pack_deep(ilub_handle* h, int slot, cudaStream_t s) {
    DeepSolve& D = h->solve[slot].deep;
    if (!D.built) return;
    const int exact = options().exact != 0 ? 1 : 0;
    k_deep_pack<<<blocks_for((int64_t)h->n * 8, 256), 256, 0, s>>>(h->n, exact ? 0 : 1, D.meta.p, D.dsrc.p, h->lu.p, D.val.p,
                                                                  D.diag.p);
    h->launches++;
    D.packed_exact = exact;
}

void run_deep(ilub_handle* h, int slot, double* x, int64_t ldx, int nrhs, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    DeepSolve& D = T.deep;
    const bool exact = options().exact != 0;
    if (D.packed_exact != (exact ? 1 : 0)) pack_deep(h, slot, s);
    const size_t smem = (size_t)kDeepRing * sizeof(double);
    static thread_local bool attr_set[2] = {false, false};
    if (!attr_set[exact ? 1 : 0]) {
        if (exact)
            ILUB_CUDA(cudaFuncSetAttribute(k_trsv_deep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
            ILUB_CUDA(cudaFuncSetAttribute(k_trsv_deep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[exact ? 1 : 0] = true;
    }
    DeepParams P;
    P.n = h->n;
    P.nlevels = T.nlevels;
    P.nent = std::max(1, D.nent);
    P.lvl_ptr = T.lvl_ptr.p;
    P.meta = D.meta.p;
    P.dpos = D.dpos.p;
    P.val = D.val.p;
    P.diag = D.diag.p;
    P.bs = D.bs.p;
    P.xs = D.xs.p;
    P.unit_diag = T.unit_diag ? 1 : 0;
    for (int c = 0; c < nrhs; ++c) {
        P.x = x + (int64_t)c * ldx;
        k_deep_permute_in<<<blocks_for(h->n, 256), 256, 0, s>>>(h->n, D.meta.p, P.x, D.bs.p);
        if (exact)
            k_trsv_deep<true><<<1, kDeepThreads, smem, s>>>(P);
        else
            k_trsv_deep<false><<<1, kDeepThreads, smem, s>>>(P);
        h->launches += 2;
    }
    ILUB_CUDA(cudaGetLastError());
}

}  // namespace ilub
