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

constexpr int kDeepG = 8;             // lanes per row
constexpr int kDeepRing = 8192;       // solved values kept in shared memory (64 KB)

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
    const int* eptr;
    const double* bs;   // right-hand side in solve order
    double* xs;         // solution in solve order
    double* x;          // solution in row order (in/out vector of the solve)
    int unit_diag;
    int* status;
};

// ---- decoupled access / execute.  STAGER warps run ahead of the level loop and do every global read: they copy
// the packed rows into shared-memory rings (per position: first entry, length, row, right-hand side (times 1/d);
// per entry: coefficient and either the VALUE of the dependency, gathered from xs through L2 because it was final
// when the entry was staged, or -- for a dependency solved too recently for that -- its position).  The COMPUTE
// warps walk the levels behind one named barrier per level and touch only shared memory: ring of staged rows, ring
// of the most recent kDeepRing solved values.  A level therefore costs a barrier, three dependent shared-memory
// reads, a few FMAs and a shuffle reduction -- no L2 round trip.
constexpr int kStagers = 6;                                  // stager warps (interleaved blocks of 32 positions)
constexpr int kComputeWarps = 8;                             // 32 groups of 8 lanes
constexpr int kPosCap = 2048, kEntCap = 4096;                // staged rows / entries (powers of two)
constexpr int kBigRow = 1024;                                // rows longer than this are not staged (read from global)
constexpr int kOldMargin = 256;                              // positions this far behind the level loop are gathered by value
struct DeepSmem {
    double xring[kDeepRing];
    double ev[kEntCap];      // coefficient
    double ex[kEntCap];      // value of the dependency (if gathered at staging time)
    int ep[kEntCap];         // position of the dependency, or -1: use ex
    double pb[kPosCap];      // right-hand side (pre-scaled by 1/d unless exact)
    double pd[kPosCap];      // diagonal (exact arithmetic only)
    int4 pm[kPosCap];        // {first entry, length, row, big-row flag}
    int staged;              // blocks of 32 positions completely staged (committed in block order)
    int c_pos;               // positions below this are solved (start of the level being resolved)
    int c_q;                 // entries below this are consumed
    int abort;
};

template <bool EXACT>
__global__ void __launch_bounds__((kComputeWarps + kStagers) * 32, 1) k_trsv_deep(const DeepParams P) {
    extern __shared__ __align__(16) unsigned char deep_smem_raw[];
    DeepSmem& S = *reinterpret_cast<DeepSmem*>(deep_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) S.staged = 0, S.c_pos = 0, S.c_q = 0, S.abort = 0;
    __syncthreads();

    if (warp >= kComputeWarps) {
        // ------------------------------------------------------------------ stager: blocks of 32 positions, block b
        // belongs to stager b % kStagers; one lane per position for the row record, then the block's entries
        // (contiguous in the packed arrays) spread over the lanes
        const int w = warp - kComputeWarps;
        const int nblk = (P.n + 31) / 32;
        int mine = 0;
#ifdef ILUB_DEEP_PROFILE
        long long s_load = 0, s_space = 0, s_ent = 0, s_pub = 0, sc = clock64();
#define ST_TICK(var) { const long long tn = clock64(); var += tn - sc; sc = tn; }
#else
#define ST_TICK(var)
#endif
        for (int blk = w; blk < nblk; blk += kStagers, ++mine) {
            const int p0 = blk * 32, pos = p0 + lane;
            const int pend = min(P.n, p0 + 32);
            int4 m = make_int4(0, 0, 0, 0);
            double b = 0.0, d = 1.0;
            if (pos < P.n) {
                m = __ldg(P.meta + pos);
                b = __ldg(P.bs + pos);
                d = __ldg(P.diag + pos);
            }
            const int q0 = __shfl_sync(0xffffffffu, m.y, 0);
            const int qe = __ldg(P.eptr + pend);
            // space: position slots are free once the level loop has passed them; entry slots likewise
            ST_TICK(s_load)
            // nothing of a block is staged if its entries, together with those of the 32 rows before it (a round of
            // the level loop may start that far back), do not fit the entry ring: only with very long rows
            const bool blockbig = qe - __ldg(P.eptr + max(0, p0 - 32)) > kEntCap;
            while (true) {
                const int cp = *(volatile int*)&S.c_pos, cq = *(volatile int*)&S.c_q;
                if (pend - cp <= kPosCap && (blockbig || qe - cq <= kEntCap)) break;
                if (*(volatile int*)&S.abort) return;
                __nanosleep(100);
            }
            ST_TICK(s_space)
            const int frontier = *(volatile int*)&S.c_pos - kOldMargin;  // positions below are final and visible in L2
            __threadfence_block();
            if (pos < P.n) {
                const bool big = blockbig || m.z > kBigRow;
                S.pm[pos & (kPosCap - 1)] = make_int4(m.y, m.z, m.x, big ? 1 : 0);
                S.pb[pos & (kPosCap - 1)] = EXACT ? b : b * d;
                if (EXACT) S.pd[pos & (kPosCap - 1)] = d;
            }
            // entries of the block (rows longer than kBigRow are skipped: the consumer reads them from global);
            // a block whose entries exceed the ring (only possible with big rows) stages nothing
            if (!blockbig) {
                // eight entries per lane in flight: first every position / coefficient, then every gather
                for (int qb = q0 + lane; qb < qe; qb += 32 * 8) {
                    int dpv[8];
                    double vv[8], xv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int q = qb + u * 32;
                        dpv[u] = -2;
                        if (q < qe) {
                            dpv[u] = __ldg(P.dpos + q);
                            vv[u] = __ldg(P.val + q);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (dpv[u] >= 0 && dpv[u] < frontier) xv[u] = __ldcg(P.xs + dpv[u]);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (dpv[u] >= 0) {
                            const int e = (qb + u * 32) & (kEntCap - 1);
                            S.ev[e] = vv[u];
                            if (dpv[u] < frontier) {
                                S.ex[e] = xv[u];
                                S.ep[e] = -1;
                            } else {
                                S.ep[e] = dpv[u];
                            }
                        }
                    }
                }
            }
            ST_TICK(s_ent)
            __syncwarp();
            __threadfence_block();
            if (lane == 0) {  // commit in block order: consumers test one counter
                while (*(volatile int*)&S.staged != blk) {
                    if (*(volatile int*)&S.abort) break;
                }
                *(volatile int*)&S.staged = blk + 1;
            }
            __syncwarp();
            ST_TICK(s_pub)
        }
#ifdef ILUB_DEEP_PROFILE
        if (lane == 0 && w < 2) printf("stager %d: blocks %d | cycles per block: row loads %.0f space wait %.0f entries %.0f publish %.0f\n", w, mine, (double)s_load / mine, (double)s_space / mine, (double)s_ent / mine, (double)s_pub / mine);
#endif
        return;
    }

    // ---------------------------------------------------------------------- compute warps
    const int gl = threadIdx.x % kDeepG;
    const int g = threadIdx.x / kDeepG;
    const int ng = kComputeWarps * 32 / kDeepG;
    const unsigned gmask = 0xffu << ((lane / kDeepG) * kDeepG);
    int lb = __ldg(P.lvl_ptr), le = __ldg(P.lvl_ptr + 1);
    int seen = 0;  // positions below this are known to be staged
#ifdef ILUB_DEEP_PROFILE
    long long t_wait = 0, t_row = 0, t_write = 0, t_bar = 0, t_pub = 0, tc = clock64();
#define DEEP_TICK(var) { const long long tn = clock64(); var += tn - tc; tc = tn; }
#else
#define DEEP_TICK(var)
#endif
    for (int l = 0; l < P.nlevels; ++l) {
        const int lne = l + 1 < P.nlevels ? __ldg(P.lvl_ptr + l + 2) : le;
        // rounds of `ng` rows: a level wider than that is resolved in several rounds, each behind its own barrier, so
        // that the position the stagers measure their space against never lags more than one round
        for (int base = lb; base < le; base += ng) {
        const int pos = base + g;
        const int rend = min(le, base + ng);
        if (pos < le) {
            // the block of this position has been staged
            if (pos >= seen) {
                unsigned spins = 0;
                while (true) {
                    seen = *(volatile int*)&S.staged * 32;
                    if (pos < seen) break;
                    if ((++spins & 0xffffu) == 0 && spins > (1u << 28)) {
                        *(volatile int*)&S.abort = 1;
                        break;
                    }
                }
                __threadfence_block();
            }
            DEEP_TICK(t_wait)
            const int4 m = S.pm[pos & (kPosCap - 1)];  // {first entry, length, row, big}
            double acc = gl == 0 ? S.pb[pos & (kPosCap - 1)] : 0.0;
            if (m.w == 0 && m.y <= kEntCap) {
                if (EXACT) {
                    if (gl == 0) {
                        for (int j = 0; j < m.y; ++j) {
                            const int e = (m.x + j) & (kEntCap - 1);
                            const int dp = S.ep[e];
                            const double xv = dp < 0 ? S.ex[e] : (dp >= rend - kDeepRing ? S.xring[dp & (kDeepRing - 1)] : __ldcg(P.xs + dp));
                            acc = __dsub_rn(acc, __dmul_rn(S.ev[e], xv));
                        }
                    }
                } else {
                    for (int j = gl; j < m.y; j += kDeepG) {
                        const int e = (m.x + j) & (kEntCap - 1);
                        const int dp = S.ep[e];
                        // one shared-memory read either way (gathered value or ring); the global fall-back only
                        // exists for levels wider than the ring
                        const double* src = dp < 0 ? &S.ex[e] : &S.xring[dp & (kDeepRing - 1)];
                        double xv = *src;
                        if (dp >= 0 && dp < rend - kDeepRing) xv = __ldcg(P.xs + dp);
                        acc = fma(-S.ev[e], xv, acc);
                    }
                }
            } else {
                // long row: straight from the packed arrays in global memory
                if (EXACT) {
                    if (gl == 0)
                        for (int j = 0; j < m.y; ++j) {
                            const int dp = __ldg(P.dpos + m.x + j);
                            const double xv = dp >= rend - kDeepRing ? S.xring[dp & (kDeepRing - 1)] : __ldcg(P.xs + dp);
                            acc = __dsub_rn(acc, __dmul_rn(__ldg(P.val + m.x + j), xv));
                        }
                } else {
                    for (int j = gl; j < m.y; j += kDeepG) {
                        const int dp = __ldg(P.dpos + m.x + j);
                        const double xv = dp >= rend - kDeepRing ? S.xring[dp & (kDeepRing - 1)] : __ldcg(P.xs + dp);
                        acc = fma(-__ldg(P.val + m.x + j), xv, acc);
                    }
                }
            }
            DEEP_TICK(t_row)
            if (EXACT) {
                if (gl == 0 && !P.unit_diag) acc = __ddiv_rn(acc, S.pd[pos & (kPosCap - 1)]);
            } else {
                acc += __shfl_xor_sync(gmask, acc, 4, kDeepG);
                acc += __shfl_xor_sync(gmask, acc, 2, kDeepG);
                acc += __shfl_xor_sync(gmask, acc, 1, kDeepG);
            }
            if (gl == 0) {
                S.xring[pos & (kDeepRing - 1)] = acc;
                __stcg(P.xs + pos, acc);
                P.x[m.z] = acc;
            }
            DEEP_TICK(t_write)
        }
        DEEP_TICK(t_wait)
        asm volatile("bar.sync 1, %0;" ::"n"(kComputeWarps * 32) : "memory");
        DEEP_TICK(t_bar)
        if (threadIdx.x == kComputeWarps * 32 - 1) {
            // every row below `rend` is solved and behind the barrier; this thread's fence (cumulative over what the
            // barrier ordered before it) makes the xs stores visible before the stagers see the new position.  It is
            // the last compute thread: its group only has rows in wide levels, so the fence is off the critical path.
            __threadfence_block();
            *(volatile int*)&S.c_pos = rend;
            *(volatile int*)&S.c_q = rend < P.n ? __ldg(P.eptr + rend) : P.nent;
        }
        if (*(volatile int*)&S.abort) {
            if (threadIdx.x == 0) atomicCAS(P.status, 0, 301);
            return;
        }
        DEEP_TICK(t_pub)
        }
        lb = le;
        le = lne;
    }
#ifdef ILUB_DEEP_PROFILE
    if (threadIdx.x == 0 || threadIdx.x == 64 || threadIdx.x == 248)
        printf("deep profile thread %3d: levels %d | cycles per level: wait %.0f row %.0f reduce+write %.0f barrier %.0f publish %.0f\n", threadIdx.x, P.nlevels,
               (double)t_wait / P.nlevels, (double)t_row / P.nlevels, (double)t_write / P.nlevels, (double)t_bar / P.nlevels, (double)t_pub / P.nlevels);
#endif
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
    const size_t smem = sizeof(DeepSmem);
    static thread_local unsigned attr_set[2] = {0, 0};
    if (first_use_on_this_device(attr_set, exact ? 1 : 0)) {
        if (exact)
            ILUB_CUDA(cudaFuncSetAttribute(k_trsv_deep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
            ILUB_CUDA(cudaFuncSetAttribute(k_trsv_deep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    DeepParams P;
    P.n = h->n;
    P.nlevels = T.nlevels;
    P.nent = std::max(1, D.nent);
    P.lvl_ptr = T.lvl_ptr.p;
    P.meta = D.meta.p;
    P.eptr = D.eptr.p;
    P.dpos = D.dpos.p;
    P.val = D.val.p;
    P.diag = D.diag.p;
    P.bs = D.bs.p;
    P.xs = D.xs.p;
    P.unit_diag = T.unit_diag ? 1 : 0;
    P.status = h->status.p;
    for (int c = 0; c < nrhs; ++c) {
        P.x = x + (int64_t)c * ldx;
        k_deep_permute_in<<<blocks_for(h->n, 256), 256, 0, s>>>(h->n, D.meta.p, P.x, D.bs.p);
        if (exact)
            k_trsv_deep<true><<<1, (kComputeWarps + kStagers) * 32, smem, s>>>(P);
        else
            k_trsv_deep<false><<<1, (kComputeWarps + kStagers) * 32, smem, s>>>(P);
        h->launches += 2;
    }
    ILUB_CUDA(cudaGetLastError());
}

}  // namespace ilub
