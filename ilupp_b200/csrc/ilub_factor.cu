// Numeric ILU(0) / IChol(0) on the pattern of A -- replaces compute_ilu0 + sparse_vec_update
// (ILU0.hpp:8-66) and compute_ichol0 + sparse_dot_product (IChol.hpp:15-59).
//
// Two schedules over the same per-row routine:
//   sync-free      one persistent launch; a group of G lanes owns a row, rows are taken in
//                  (CTA part, level, row) order -- the order of the L solve -- and a row waits for
//                  the rows of its strictly-lower columns through acquire/release epoch flags;
//   level-scheduled one launch per level of the L DAG, no flags.
// The reference's two-pointer merges become lane-parallel intersections: every lane of the group
// takes entries of row i and binary-searches their column in row k (ILU0) / row j (IChol0).
// Operation order per entry is the reference's (updates in ascending k; IChol0 dot product summed
// in ascending column order via an ordered shuffle reduction), so with `exact` arithmetic the
// factors are bit-for-bit those of the reference compiled without FMA contraction.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "ilub_common.cuh"
#include "ilub_ptx.cuh"

namespace ilub {
namespace {

struct FactorParams {
    int n;
    const int* ptr;
    const int* idx;
    const int* dpos;
    const double* a;
    double* lu;
    // sync-free schedule
    const int* part_chunk_ptr;
    const int* rows;  // chunk-lane -> row (-1 padding)
    unsigned char* fflag;
    unsigned epoch;
    // level schedule
    const int* lvl_rows;
    int lvl_begin, lvl_end;
    int* status;
    long long spin_limit;
    // level-scheduled kernels: one big row buffer per CTA (28 bytes x big_cap, behind the groups' buffers, then a lock
    // word) for the rare rows longer than 4 * G entries -- hub rows of irregular matrices, which the slow row routines
    // would otherwise turn into milliseconds on the critical path
    int big_cap;
    // level loop on deep narrow DAGs: the last `nprefetch` CTAs of the grid do not factorise; they run a few levels ahead
    // of the working CTAs and pull what those levels will read into L2 (see k_factor_levels)
    int nprefetch;
};

// first position in [b, e) with idx >= c
__device__ __forceinline__ int lower_bound(const int* __restrict__ idx, int b, int e, int c) {
    while (b < e) {
        const int m = (b + e) >> 1;
        if (__ldg(idx + m) < c)
            b = m + 1;
        else
            e = m;
    }
    return b;
}

template <bool SYNCFREE>
__device__ __forceinline__ bool wait_row(const FactorParams& P, int k, long long& t0) {
    if (!SYNCFREE) return true;
    unsigned spins = 0;
    while (ld_acquire_gpu_u8(P.fflag + k) != P.epoch) {
        if ((++spins & 63u) == 0) {
            if (t0 == 0) t0 = clock64();
            if (clock64() - t0 > P.spin_limit || *(volatile int*)P.status != 0) {
                atomicCAS(P.status, 0, 101);
                return false;
            }
        }
    }
    return true;
}

// ILU(0) of one row by a group of G lanes (gl = lane in group, gmask = the group's lanes)
template <int G, bool SYNCFREE, bool EXACT>
__device__ __forceinline__ bool ilu0_row(const FactorParams& P, int i, int gl, unsigned gmask) {
    const int p0 = P.ptr[i], p1 = P.ptr[i + 1], dp = P.dpos[i];
    for (int p = p0 + gl; p < p1; p += G) P.lu[p] = P.a[p];  // ILU0.hpp:37
    long long t0 = 0;
    for (int kk = p0; kk < dp; ++kk) {  // ILU0.hpp:47-62, ascending k
        const int k = __ldg(P.idx + kk);
        if (!wait_row<SYNCFREE>(P, k, t0)) return false;
        const int dk = P.dpos[k], ek = P.ptr[k + 1];
        const double ukk = __ldcg(P.lu + dk);
        const int owner = (kk - p0) % G;
        double mine = 0.0;
        if (gl == owner) mine = P.lu[kk];
        const double wkk = __shfl_sync(gmask, mine, owner, G);
        const double lik = EXACT ? __ddiv_rn(wkk, ukk) : wkk / ukk;  // ILU0.hpp:52
        // U[i, k+1:] -= L_ik * U[k, k+1:]   (ILU0.hpp:55-59): my entries right of kk
        int p = p0 + gl;
        if (p <= kk) p += ((kk - p) / G + 1) * G;
        for (; p < p1; p += G) {
            const int c = __ldg(P.idx + p);
            const int q = lower_bound(P.idx, dk + 1, ek, c);
            if (q < ek && __ldg(P.idx + q) == c) P.lu[p] = sub_mul<EXACT>(P.lu[p], lik, __ldcg(P.lu + q));
        }
        if (gl == owner) P.lu[kk] = lik;  // ILU0.hpp:61
    }
    return true;
}

// IChol(0) of one row: entries [p0, dp] of the CSR-read arrays are row i of tril(A), diagonal last
template <int G, bool SYNCFREE, bool EXACT>
__device__ __forceinline__ bool ichol0_row(const FactorParams& P, int i, int gl, unsigned gmask, int gshift) {
    const int p0 = P.ptr[i], dp = P.dpos[i];
    long long t0 = 0;
    for (int k = p0; k <= dp; ++k) {  // IChol.hpp:39-57
        const int j = __ldg(P.idx + k);
        int jb = p0, je = dp;
        double ljj = 0.0;
        if (j < i) {
            if (!wait_row<SYNCFREE>(P, j, t0)) return false;
            jb = P.ptr[j];
            je = P.dpos[j];
            ljj = __ldcg(P.lu + je);  // diagonal is the last entry of row j of L (IChol.hpp:50)
        }
        // dp = sum over common columns of (row i \ diag) and (row j \ diag), ascending column
        double sum = 0.0;
        for (int pb = p0; pb < k; pb += G) {
            const int p = pb + gl;
            double prod = 0.0;
            bool hit = false;
            if (p < k) {
                if (j == i) {
                    const double v = P.lu[p];
                    prod = EXACT ? __dmul_rn(v, v) : v * v;
                    hit = true;
                } else {
                    const int c = __ldg(P.idx + p);
                    const int q = lower_bound(P.idx, jb, je, c);
                    if (q < je && __ldg(P.idx + q) == c) {
                        const double v = P.lu[p], u = __ldcg(P.lu + q);
                        prod = EXACT ? __dmul_rn(v, u) : v * u;
                        hit = true;
                    }
                }
            }
            unsigned m = (__ballot_sync(gmask, hit) & gmask) >> gshift;
            while (m) {  // ordered accumulation, IChol.hpp:18-27
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const double t = __shfl_sync(gmask, prod, src, G);
                sum = EXACT ? __dadd_rn(sum, t) : sum + t;
            }
        }
        const double aij = P.a[k];
        double v;
        if (j < i)
            v = EXACT ? __ddiv_rn(__dsub_rn(aij, sum), ljj) : (aij - sum) / ljj;  // IChol.hpp:49-51
        else
            v = EXACT ? __dsqrt_rn(__dsub_rn(aij, sum)) : sqrt(aij - sum);        // IChol.hpp:52-53
        if (gl == (k - p0) % G) P.lu[k] = v;
        __syncwarp(gmask);
    }
    return true;
}


// ------------------------------------------------------------------------------------------------
// Short rows (the common case: at most 4 * G entries) are factorised in SHARED memory.  The row routines above
// spend their time in chains of dependent global loads (per strictly-lower entry: column -> diagonal position ->
// pivot, then a binary search through row k in L2: ~8 round trips, ~25 per stencil row).  Here a row costs
//   1 round trip  to stage row i (indices, values),
//   3 round trips to fetch {diagonal position, row end, pivot} of ALL its strictly-lower columns in parallel,
//   1 round trip  per strictly-lower entry for row k's upper part (coalesced), matched against row i by a binary
//                 search in shared memory.
// The arithmetic and its order are those of the routines above (ascending k; IChol0 dot products accumulated in
// ascending column order), so `exact` results stay bit-for-bit the reference's.
template <int G>
struct RowSmem {
    static constexpr int L = 4 * G;        // longest row handled here
    static constexpr int kBytes = 28 * L;  // per group
    double* val;   // [L] row i values (become the factor's)
    double* piv;   // [L] per strictly-lower entry: pivot u_kk (ILU0) / l_jj (IChol0)
    int* idx;      // [L] row i columns
    int* ra;       // [L] ILU0: first upper entry of row k / IChol0: first entry of row j
    int* rb;       // [L] ILU0: end of row k / IChol0: diagonal position of row j
    __device__ RowSmem(unsigned char* base, int cap = L) {
        val = reinterpret_cast<double*>(base);
        piv = val + cap;
        idx = reinterpret_cast<int*>(piv + cap);
        ra = idx + cap;
        rb = ra + cap;
    }
};

__device__ __forceinline__ int lower_bound_s(const int* s, int b, int e, int c) {
    while (b < e) {
        const int m = (b + e) >> 1;
        if (s[m] < c)
            b = m + 1;
        else
            e = m;
    }
    return b;
}

// kb / ke: the strictly-lower entries [kb, ke) of the row are eliminated (ke < 0: all of them); from_lu: the row's current
// values are taken from the factor array (an earlier call has eliminated [0, kb)) instead of from A.  The elimination
// runs in ascending column order whatever the split, so splitting a row does not change a single operation.
// do_stage / do_write / meta_to: a row may be eliminated in two calls that keep it in the group's buffer in between -- the
// first call stages it, fetches the static {first upper entry, row end} of ALL rows up to meta_to and does not write the row
// back; the second neither stages nor looks the extents up again, only the pivots of its own range.
template <int G, bool SYNCFREE, bool EXACT>
__device__ __forceinline__ bool ilu0_row_fast(const FactorParams& P, int i, int gl, unsigned gmask, const RowSmem<G>& R,
                                              int kb = 0, int ke = -1, bool from_lu = false, bool do_stage = true,
                                              bool do_write = true, int meta_to = -1) {
    const int p0 = P.ptr[i], p1 = P.ptr[i + 1], dp = P.dpos[i];
    const int len = p1 - p0, nl = ke < 0 ? dp - p0 : ke;
    if (do_stage) {
        for (int t = gl; t < len; t += G) {  // ILU0.hpp:37
            R.idx[t] = __ldg(P.idx + p0 + t);
            R.val[t] = from_lu ? __ldcg(P.lu + p0 + t) : P.a[p0 + t];
        }
        __syncwarp(gmask);
    }
    long long t0 = 0;
    bool ok = true;
    for (int t = kb + gl; t < max(nl, meta_to); t += G) {
        const int k = R.idx[t];
        if (!wait_row<SYNCFREE>(P, k, t0)) ok = false;
        if (do_stage) {
            R.ra[t] = P.dpos[k] + 1;
            R.rb[t] = P.ptr[k + 1];
        }
        if (t < nl) R.piv[t] = __ldcg(P.lu + R.ra[t] - 1);
    }
    if (SYNCFREE && !__all_sync(gmask, ok)) return false;
    __syncwarp(gmask);
    // row k's upper parts are final and independent of the updates below: the first G entries of the next four rows
    // are requested ahead, so the loop runs at the pace of the shared-memory work, not of one L2 round trip per k
    constexpr int PF = 4;
    int pc[PF];
    double pu[PF];
    auto fetch = [&](int kk, int& c, double& u) {
        c = -1;
        if (kk < nl) {
            const int q = R.ra[kk] + gl;
            if (q < R.rb[kk]) {
                c = __ldg(P.idx + q);
                u = __ldcg(P.lu + q);
            }
        }
    };
#pragma unroll
    for (int u = 0; u < PF; ++u) fetch(kb + u, pc[u], pu[u]);
    for (int kk0 = kb; kk0 < nl; kk0 += PF) {  // ILU0.hpp:47-62, ascending k
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int kk = kk0 + u;
            if (kk < nl) {
                const double wkk = R.val[kk], ukk = R.piv[kk];
                const double lik = EXACT ? __ddiv_rn(wkk, ukk) : wkk / ukk;  // ILU0.hpp:52
                const int qb = R.ra[kk], qe = R.rb[kk];
                // U[i, k+1:] -= L_ik * U[k, k+1:]   (ILU0.hpp:55-59): row k's upper part against my entries right of kk
                if (pc[u] >= 0) {
                    const int pos = lower_bound_s(R.idx, kk + 1, len, pc[u]);
                    if (pos < len && R.idx[pos] == pc[u]) R.val[pos] = sub_mul<EXACT>(R.val[pos], lik, pu[u]);
                }
                for (int q = qb + gl + G; q < qe; q += G) {
                    const int c = __ldg(P.idx + q);
                    const double uu = __ldcg(P.lu + q);
                    const int pos = lower_bound_s(R.idx, kk + 1, len, c);
                    if (pos < len && R.idx[pos] == c) R.val[pos] = sub_mul<EXACT>(R.val[pos], lik, uu);
                }
                fetch(kk + PF, pc[u], pu[u]);
                __syncwarp(gmask);
                if (gl == 0) R.val[kk] = lik;  // ILU0.hpp:61
            }
        }
    }
    __syncwarp(gmask);
    if (do_write)
        for (int t = gl; t < len; t += G) P.lu[p0 + t] = R.val[t];
    return true;
}

template <int G, bool SYNCFREE, bool EXACT>
__device__ __forceinline__ bool ichol0_row_fast(const FactorParams& P, int i, int gl, unsigned gmask, int gshift,
                                                const RowSmem<G>& R) {
    const int p0 = P.ptr[i], dp = P.dpos[i];
    const int len = dp - p0 + 1;  // row i of tril(A), diagonal last
    for (int t = gl; t < len; t += G) {
        R.idx[t] = __ldg(P.idx + p0 + t);
        R.val[t] = P.a[p0 + t];
    }
    __syncwarp(gmask);
    long long t0 = 0;
    bool ok = true;
    for (int t = gl; t < len - 1; t += G) {
        const int j = R.idx[t];
        if (!wait_row<SYNCFREE>(P, j, t0)) ok = false;
        const int je = P.dpos[j];
        R.ra[t] = P.ptr[j];
        R.rb[t] = je;
        R.piv[t] = __ldcg(P.lu + je);  // diagonal is the last entry of row j of L (IChol.hpp:50)
    }
    if (SYNCFREE && !__all_sync(gmask, ok)) return false;
    __syncwarp(gmask);
    for (int t = 0; t < len; ++t) {  // IChol.hpp:39-57
        const bool diag = t == len - 1;
        const int qb = diag ? 0 : R.ra[t], qe = diag ? t : R.rb[t];
        // sum over the common columns of (row i left of t) and (row j without its diagonal), ascending column
        double sum = 0.0;
        for (int q0 = qb; q0 < qe; q0 += G) {
            const int q = q0 + gl;
            double prod = 0.0;
            bool hit = false;
            if (q < qe) {
                if (diag) {
                    const double v = R.val[q];
                    prod = EXACT ? __dmul_rn(v, v) : v * v;
                    hit = true;
                } else {
                    const int c = __ldg(P.idx + q);
                    const int pos = lower_bound_s(R.idx, 0, t, c);
                    if (pos < t && R.idx[pos] == c) {
                        const double v = R.val[pos], u = __ldcg(P.lu + q);
                        prod = EXACT ? __dmul_rn(v, u) : v * u;
                        hit = true;
                    }
                }
            }
            unsigned m = (__ballot_sync(gmask, hit) & gmask) >> gshift;
            while (m) {  // ordered accumulation, IChol.hpp:18-27
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const double tt = __shfl_sync(gmask, prod, src, G);
                sum = EXACT ? __dadd_rn(sum, tt) : sum + tt;
            }
        }
        const double aij = R.val[t];
        double v;
        if (!diag) {
            const double ljj = R.piv[t];
            v = EXACT ? __ddiv_rn(__dsub_rn(aij, sum), ljj) : (aij - sum) / ljj;  // IChol.hpp:49-51
        } else {
            v = EXACT ? __dsqrt_rn(__dsub_rn(aij, sum)) : sqrt(aij - sum);        // IChol.hpp:52-53
        }
        __syncwarp(gmask);
        if (gl == 0) R.val[t] = v;
        __syncwarp(gmask);
    }
    for (int t = gl; t < len; t += G) P.lu[p0 + t] = R.val[t];
    return true;
}

// one row: short (the group's own buffer), long (the CTA's big buffer, one long row at a time), or longer still
template <int G, bool ICHOL, bool SYNCFREE, bool EXACT>
__device__ __forceinline__ bool factor_row(const FactorParams& P, int i, int gl, unsigned gmask, int gshift,
                                           const RowSmem<G>& R, unsigned char* big) {
    const int len = ICHOL ? P.dpos[i] - P.ptr[i] + 1 : P.ptr[i + 1] - P.ptr[i];
    if (len <= RowSmem<G>::L) {
        if (ICHOL) return ichol0_row_fast<G, SYNCFREE, EXACT>(P, i, gl, gmask, gshift, R);
        return ilu0_row_fast<G, SYNCFREE, EXACT>(P, i, gl, gmask, R);
    }
    if (!SYNCFREE && len <= P.big_cap) {  // (never while rows wait on each other: the lock holder could wait forever)
        int* lock = reinterpret_cast<int*>(big + (size_t)28 * P.big_cap);
        if (gl == 0)
            while (atomicCAS(lock, 0, 1) != 0) __nanosleep(64);
        __syncwarp(gmask);
        const RowSmem<G> RB(big, P.big_cap);
        bool ok;
        if (ICHOL) ok = ichol0_row_fast<G, SYNCFREE, EXACT>(P, i, gl, gmask, gshift, RB);
        else ok = ilu0_row_fast<G, SYNCFREE, EXACT>(P, i, gl, gmask, RB);
        __syncwarp(gmask);
        if (gl == 0) {
            __threadfence_block();
            atomicExch(lock, 0);
        }
        return ok;
    }
    if (ICHOL) return ichol0_row<G, SYNCFREE, EXACT>(P, i, gl, gmask, gshift);
    return ilu0_row<G, SYNCFREE, EXACT>(P, i, gl, gmask);
}

template <int G, bool ICHOL, bool SYNCFREE, bool EXACT>
__global__ void __launch_bounds__(512) k_factor(FactorParams P) {
    extern __shared__ __align__(16) unsigned char factor_smem[];
    const RowSmem<G> R(factor_smem + (size_t)(threadIdx.x / G) * RowSmem<G>::kBytes);
    unsigned char* big = factor_smem + (size_t)(blockDim.x / G) * RowSmem<G>::kBytes;
    if (!SYNCFREE && P.big_cap > 0) {
        if (threadIdx.x == 0) *reinterpret_cast<int*>(big + (size_t)28 * P.big_cap) = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const int gshift = (lane / G) * G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gshift);
    const int gidx = threadIdx.x / G;
    const int ng = blockDim.x / G;
    if (SYNCFREE) {
        const int64_t s0 = (int64_t)P.part_chunk_ptr[blockIdx.x] * 32;
        const int64_t s1 = (int64_t)P.part_chunk_ptr[blockIdx.x + 1] * 32;
        for (int64_t s = s0 + gidx; s < s1; s += ng) {
            const int i = P.rows[s];
            if (i < 0) continue;
            if (!factor_row<G, ICHOL, true, EXACT>(P, i, gl, gmask, gshift, R, big)) return;
            __syncwarp(gmask);
            if (gl == 0) {
                __threadfence();
                st_relaxed_gpu_u8(P.fflag + i, P.epoch);
            }
        }
    } else {
        const int64_t s = (int64_t)P.lvl_begin + (int64_t)blockIdx.x * ng + gidx;
        if (s < P.lvl_end) {
            const int i = P.lvl_rows[s];
            factor_row<G, ICHOL, false, EXACT>(P, i, gl, gmask, gshift, R, big);
        }
    }
}

// Level loop: ONE cooperative launch walks all levels of the L DAG; between two levels every CTA passes a
// grid-wide barrier (a monotone arrival counter in global memory).  Right when levels are few and wide
// (stencil matrices, block-diagonal batches): ~2 us per level instead of a dependency flag per row.
template <int G, bool ICHOL, bool EXACT>
__global__ void __launch_bounds__(512) k_factor_levels(FactorParams P, const int* __restrict__ lvl_ptr, int nlevels,
                                                       unsigned* barrier) {
    extern __shared__ __align__(16) unsigned char factor_smem[];
    const RowSmem<G> R(factor_smem + (size_t)(threadIdx.x / G) * RowSmem<G>::kBytes);
    unsigned char* big = factor_smem + (size_t)(blockDim.x / G) * RowSmem<G>::kBytes;
    if (P.big_cap > 0) {
        if (threadIdx.x == 0) *reinterpret_cast<int*>(big + (size_t)28 * P.big_cap) = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const int gshift = (lane / G) * G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gshift);
    const int ng = blockDim.x / G;
    long long t0 = 0;
    const int nwork = (int)gridDim.x - P.nprefetch;
    if ((int)blockIdx.x >= nwork) {
        // ---- prefetch CTA.  On a deep narrow DAG a level is ~9 dependent memory round trips of a handful of rows, and
        // the matrix (hundreds of MB) does not live in L2: every one of them is an HBM miss.  Nothing about WHERE a row
        // will read is dynamic, so these warps walk the rows of the levels 2 .. 2 + kAhead ahead of the workers and
        // touch row i itself, {diagonal position, row end} of its strictly-lower columns and the upper parts (indices,
        // values) of those rows with prefetch.global.L2 -- a pure hint (the values are re-read coherently from L2 when
        // they are needed), which turns the workers' misses into L2 hits.
        constexpr int kAhead = 24, kMaxRowsPerLevel = 512;
        const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        const int pw = ((int)blockIdx.x - nwork) * nw + warp, pstride = P.nprefetch * nw;
        int done = 0;
        unsigned spins = 0;
        while (true) {
            const int cur = (int)(ld_relaxed_gpu_u32(barrier) / (unsigned)nwork);
            if (cur >= nlevels - 1 || *(volatile int*)P.status != 0) break;
            const int lo = max(done, cur + 2), hi = min(nlevels, cur + 2 + kAhead);
            if (lo >= hi) {
                __nanosleep(200);
                if ((++spins & 0xffffu) == 0) {
                    if (t0 == 0) t0 = clock64();
                    if (clock64() - t0 > P.spin_limit) break;
                }
                continue;
            }
            for (int l = lo; l < hi; ++l) {
                const int b = __ldg(lvl_ptr + l), e = min(__ldg(lvl_ptr + l + 1), b + kMaxRowsPerLevel);
                for (int s = b + pw; s < e; s += pstride) {
                    const int i = P.lvl_rows[s];
                    const int p0 = P.ptr[i], dp = P.dpos[i], pe = P.ptr[i + 1];
                    for (int t = p0 + lane * 16; t < pe; t += 32 * 16) {  // row i: 128 bytes of values per line
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.a + t));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.idx + t));
                    }
                    for (int t = p0 + lane; t < dp; t += 32) {  // rows k of the strictly-lower columns
                        const int k = P.idx[t];
                        const int dk = P.dpos[k], ek = P.ptr[k + 1];
                        for (int q = dk; q < ek; q += 16) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.lu + q));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.idx + q));
                        }
                        if (ek > dk) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.lu + ek - 1));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.idx + ek - 1));
                        }
                    }
                }
            }
            done = hi;
        }
        return;
    }
    const int64_t g = (int64_t)blockIdx.x * ng + threadIdx.x / G;
    const int64_t gstride = (int64_t)nwork * ng;
    for (int l = 0; l < nlevels; ++l) {
        const int b = __ldg(lvl_ptr + l), e = __ldg(lvl_ptr + l + 1);
        for (int64_t s = b + g; s < e; s += gstride) {
            const int i = P.lvl_rows[s];
            factor_row<G, ICHOL, false, EXACT>(P, i, gl, gmask, gshift, R, big);
        }
        if (l + 1 == nlevels) break;
        // grid barrier: the rows of level l are visible (lu is read through L2) before anyone starts level l + 1
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(barrier, 1u);
            const unsigned target = (unsigned)(l + 1) * (unsigned)nwork;
            unsigned spins = 0;
            while (ld_relaxed_gpu_u32(barrier) < target) {
                if ((++spins & 1023u) == 0) {
                    if (t0 == 0) t0 = clock64();
                    if (clock64() - t0 > P.spin_limit || *(volatile int*)P.status != 0) {
                        atomicCAS(P.status, 0, 102);
                        break;
                    }
                }
            }
            __threadfence();  // acquire side of the barrier (the polls above are relaxed: no fence per poll)
        }
        __syncthreads();
    }
}

template <int G, bool ICHOL>
void launch_levels_g(bool exact, cudaStream_t s, FactorParams P, const int* lvl_ptr, int nlevels, unsigned* barrier,
                     int sm_count, int64_t nrows) {
    void* kern = exact ? (void*)k_factor_levels<G, ICHOL, true> : (void*)k_factor_levels<G, ICHOL, false>;
    // few, fat CTAs: the barrier costs one atomic and one polling thread per CTA
    const int block = 512;
    const size_t smem = (size_t)(block / G) * RowSmem<G>::kBytes + (P.big_cap > 0 ? (size_t)28 * P.big_cap + 16 : 0);
    ILUB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    ILUB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
    ILUB_REQUIRE(occ > 0, ILUB_ERR_CUDA, "level-loop factorisation kernel does not fit on an SM");
    int grid = std::min(occ, 2) * sm_count;  // (measured: two CTAs per SM beat one; the level time is row latency, not the barrier)
    // deep narrow DAGs (a handful of rows per level): every CTA that joins the barrier makes the level longer and only
    // a few row groups have work -- the grid is sized to twice an average level (irregular power-law matrix, n = 10^6,
    // 9 rows per level: ONE CTA, whose barrier is a __syncthreads and an uncontended atomic)
    const int64_t avg = (nrows + nlevels - 1) / std::max(1, nlevels);
    const int ng = block / G;
    grid = (int)std::max<int64_t>(1, std::min<int64_t>(grid, (2 * avg + ng - 1) / ng));
    // ... and when that leaves most of the chip idle and the levels are many, two more CTAs prefetch ahead of the workers
    P.nprefetch = (grid <= 4 && nlevels > 1024) ? 2 : 0;
    grid += P.nprefetch;
    void* args[] = {(void*)&P, (void*)&lvl_ptr, (void*)&nlevels, (void*)&barrier};
    ILUB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(block), args, smem, s));
}

template <bool ICHOL>
void launch_levels(int G, bool exact, cudaStream_t s, const FactorParams& P, const int* lvl_ptr, int nlevels,
                   unsigned* barrier, int sm_count, int64_t nrows) {
    switch (G) {
        case 4: launch_levels_g<4, ICHOL>(exact, s, P, lvl_ptr, nlevels, barrier, sm_count, nrows); break;
        case 8: launch_levels_g<8, ICHOL>(exact, s, P, lvl_ptr, nlevels, barrier, sm_count, nrows); break;
        case 16: launch_levels_g<16, ICHOL>(exact, s, P, lvl_ptr, nlevels, barrier, sm_count, nrows); break;
        default: launch_levels_g<32, ICHOL>(exact, s, P, lvl_ptr, nlevels, barrier, sm_count, nrows); break;
    }
}

template <int G, bool ICHOL, bool SYNCFREE>
void launch_g(bool exact, int grid, int block, cudaStream_t s, const FactorParams& P) {
    const size_t smem = (size_t)(block / G) * RowSmem<G>::kBytes  // 112 bytes per thread
                        + (!SYNCFREE && P.big_cap > 0 ? (size_t)28 * P.big_cap + 16 : 0);
    static thread_local unsigned attr_set[2] = {0, 0};  // (per template instantiation)
    if (smem > 48 * 1024 && first_use_on_this_device(attr_set, exact ? 1 : 0)) {
        if (exact)
            ILUB_CUDA(cudaFuncSetAttribute(k_factor<G, ICHOL, SYNCFREE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        else
            ILUB_CUDA(cudaFuncSetAttribute(k_factor<G, ICHOL, SYNCFREE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    if (SYNCFREE) {
        // the sync-free schedule only makes progress when every CTA is resident (rows wait for rows of other CTAs)
        int occ = 0, dev = 0, sms = 0;
        if (exact)
            ILUB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_factor<G, ICHOL, SYNCFREE, true>, block, smem));
        else
            ILUB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_factor<G, ICHOL, SYNCFREE, false>, block, smem));
        ILUB_CUDA(cudaGetDevice(&dev));
        ILUB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        ILUB_REQUIRE((int64_t)occ * sms >= grid, ILUB_ERR_CUDA,
                     "sync-free factorisation: the grid cannot be co-resident on this device (lower nparts / ctas_per_sm, or use factor_algo 1 or 3)");
    }
    if (exact)
        k_factor<G, ICHOL, SYNCFREE, true><<<grid, block, smem, s>>>(P);
    else
        k_factor<G, ICHOL, SYNCFREE, false><<<grid, block, smem, s>>>(P);
}

template <bool ICHOL, bool SYNCFREE>
void launch(int G, bool exact, int grid, int block, cudaStream_t s, const FactorParams& P) {
    switch (G) {
        case 4: launch_g<4, ICHOL, SYNCFREE>(exact, grid, block, s, P); break;
        case 8: launch_g<8, ICHOL, SYNCFREE>(exact, grid, block, s, P); break;
        case 16: launch_g<16, ICHOL, SYNCFREE>(exact, grid, block, s, P); break;
        default: launch_g<32, ICHOL, SYNCFREE>(exact, grid, block, s, P); break;
    }
}

}  // namespace

__global__ void k_max_row_len(int n, const int* __restrict__ ptr, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int v = i < n ? ptr[i + 1] - ptr[i] : 0;
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    // (one look at the running maximum first: after the first few warps hardly anybody has to touch the contended word)
    if ((threadIdx.x & 31) == 0 && v > *(volatile int*)out) atomicMax(out, v);
}

// ------------------------------------------------------------------------------------------------
// Block-sequential ILU(0) for deep, narrow, band-local DAGs (BASELINE config 4) -- the factorisation counterpart of
// ilub_block.cu.  The level loop pays ~23 us per level there (a handful of rows, each a chain of dependent fetches of the
// rows it eliminates with).  Rows are taken in natural order in blocks of 32 (the blocks of the solve's layout):
//   * the eliminations of a row with FAR rows k (more than kFarWin rows back: the random long-range half of its lower
//     part) only need rows that are final long before the row's block is reached: the groups of the STAGER CTAs take one
//     block each, up to kFarWin / 32 blocks ahead of the front, and eliminate [0, nfar) of every row of their block --
//     all the HBM latency of those fetches is off the critical path;
//   * ONE CTA walks the blocks in order and finishes the rows: the remaining (near) eliminations [nfar, nl), level by
//     level of the block's local DAG (rows of one local level in parallel, one group each, __syncthreads between levels).
// Every row still sees its eliminations in ascending column order, so the result is the level loop's bit for bit.
constexpr int kFarWin = 8192;

template <int G, bool EXACT>
__device__ __forceinline__ void factor_row_range(const FactorParams& P, int i, int gl, unsigned gmask, const RowSmem<G>& R,
                                                 unsigned char* big, int kb, int ke, bool from_lu, bool do_stage = true,
                                                 bool do_write = true, int meta_to = -1) {
    const int len = P.ptr[i + 1] - P.ptr[i];
    if (len <= RowSmem<G>::L) {
        ilu0_row_fast<G, false, EXACT>(P, i, gl, gmask, R, kb, ke, from_lu, do_stage, do_write, meta_to);
        return;
    }
    int* lock = reinterpret_cast<int*>(big + (size_t)28 * P.big_cap);  // (the caller guarantees len <= big_cap)
    if (gl == 0)
        while (atomicCAS(lock, 0, 1) != 0) __nanosleep(64);
    __syncwarp(gmask);
    const RowSmem<G> RB(big, P.big_cap);
    ilu0_row_fast<G, false, EXACT>(P, i, gl, gmask, RB, kb, ke, from_lu);
    __syncwarp(gmask);
    if (gl == 0) {
        __threadfence_block();
        atomicExch(lock, 0);
    }
}

// Second step of a row of the block factorisation that stayed in its group's buffer (see k_factor_block): the
// eliminations [kb, ke) are all with rows of the SAME block, and every such row of at most RowSmem<G>::L entries still sits,
// final, in the buffer of the group that owns it (group = row - r0): pivot and upper part are read from shared memory --
// no global round trip on the critical path of a local level.  (Longer rows of the block are read from the factor array.)
template <int G, bool EXACT>
__device__ __forceinline__ void ilu0_row_local(const FactorParams& P, int i, int r0, int gl, unsigned gmask, const RowSmem<G>& R,
                                               unsigned char* smem_base, int kb, int ke) {
    const int p0 = P.ptr[i], len = P.ptr[i + 1] - p0;
    for (int kk = kb; kk < ke; ++kk) {  // ILU0.hpp:47-62, ascending k
        const int k = R.idx[kk];
        const int pk = __ldg(P.ptr + k), qb = R.ra[kk], qe = R.rb[kk];  // row k: [pk, qe), upper part [qb, qe)
        const bool in_smem = qe - pk <= RowSmem<G>::L;
        const RowSmem<G> RK(smem_base + (size_t)(k - r0) * RowSmem<G>::kBytes);
        const double wkk = R.val[kk];
        const double ukk = in_smem ? RK.val[qb - 1 - pk] : __ldcg(P.lu + qb - 1);
        const double lik = EXACT ? __ddiv_rn(wkk, ukk) : wkk / ukk;  // ILU0.hpp:52
        for (int q = qb + gl; q < qe; q += G) {  // U[i, k+1:] -= L_ik * U[k, k+1:]   (ILU0.hpp:55-59)
            const int c = in_smem ? RK.idx[q - pk] : __ldg(P.idx + q);
            const double uu = in_smem ? RK.val[q - pk] : __ldcg(P.lu + q);
            const int pos = lower_bound_s(R.idx, kk + 1, len, c);
            if (pos < len && R.idx[pos] == c) R.val[pos] = sub_mul<EXACT>(R.val[pos], lik, uu);
        }
        __syncwarp(gmask);
        if (gl == 0) R.val[kk] = lik;  // ILU0.hpp:61
    }
    __syncwarp(gmask);
    for (int t = gl; t < len; t += G) P.lu[p0 + t] = R.val[t];
}

// number of strictly-lower entries of row i whose column is below `limit` (they come first: columns ascend)
__device__ __forceinline__ int count_far(const FactorParams& P, int i, int limit) {
    int b = P.ptr[i], e = P.dpos[i];
    const int b0 = b;
    while (b < e) {
        const int m = (b + e) >> 1;
        if (__ldg(P.idx + m) < limit) b = m + 1;
        else e = m;
    }
    return b - b0;
}

// far eliminations per row (static: pattern only), computed once per handle
__global__ void k_factor_count_far(FactorParams P, int* __restrict__ nfar) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const int limit = (i / 32) * 32 - kFarWin;
    nfar[i] = limit > 0 ? count_far(P, i, limit) : 0;
}

template <int G, bool EXACT>
__global__ void __launch_bounds__(512, 1) k_factor_block(FactorParams P, int nblocks, const int* __restrict__ llev,
                                                      const int* __restrict__ bdepth, const int* __restrict__ nfar_of,
                                                      const int* __restrict__ nptr, const int* __restrict__ crec,
                                                      unsigned* far_done /*[nblocks]*/,
                                                      unsigned* progress, unsigned* next_block, unsigned epoch) {
    extern __shared__ __align__(16) unsigned char factor_smem[];
    __shared__ int abort_flag;
    const RowSmem<G> R(factor_smem + (size_t)(threadIdx.x / G) * RowSmem<G>::kBytes);
    unsigned char* big = factor_smem + (size_t)(blockDim.x / G) * RowSmem<G>::kBytes;
    if (P.big_cap > 0) {
        if (threadIdx.x == 0) *reinterpret_cast<int*>(big + (size_t)28 * P.big_cap) = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const int gshift = (lane / G) * G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gshift);
    const int grp = threadIdx.x / G, ng = blockDim.x / G;
    long long t0 = 0;
    if (blockIdx.x > 0) {
        // ---- stager groups: one block of rows at a time, the far eliminations of each of its rows
        while (true) {
            unsigned K = 0;
            if (gl == 0) K = atomicAdd(next_block, 1u);
            K = __shfl_sync(gmask, K, gshift);
            if ((int)K >= nblocks) break;
            const int r0 = (int)K * 32;
            const int limit = r0 - kFarWin;  // rows below `limit` are final once the front has passed block K - kFarWin / 32
            if (limit > 0) {
                const unsigned need = (unsigned)(limit + 31) / 32u;
                unsigned spins = 0;
                while (ld_relaxed_gpu_u32(progress) < need) {
                    __nanosleep(500);
                    if ((++spins & 0x3ffu) == 0) {
                        if (t0 == 0) t0 = clock64();
                        if (clock64() - t0 > P.spin_limit || *(volatile int*)P.status != 0) {
                            atomicCAS(P.status, 0, 103);
                            return;
                        }
                    }
                }
                __threadfence();
                for (int i = r0; i < min(P.n, r0 + 32); ++i) {
                    const int nfar = __ldg(nfar_of + i);
                    if (nfar > 0) factor_row_range<G, EXACT>(P, i, gl, gmask, R, big, 0, nfar, false);
                }
            }
            __syncwarp(gmask);
            __threadfence();
            if (gl == 0) st_relaxed_gpu_u32(far_done + K, epoch);
        }
        return;
    }
    // ---- the front: blocks in order, the near eliminations level by level of the block's local DAG
    long long tw = 0, tl = 0, tbig = 0, nbig = 0, nlev_sum = 0, tstep1 = 0;  // (debug, P.epoch != 0: where the front's time goes)
    for (int K = 0; K < nblocks; ++K) {
        long long c0 = P.epoch ? clock64() : 0;
        if (threadIdx.x == 0) {
            unsigned spins = 0;
            while (ld_relaxed_gpu_u32(far_done + K) != epoch) {
                if ((++spins & 0x3ffu) == 0) {
                    if (t0 == 0) t0 = clock64();
                    if (clock64() - t0 > P.spin_limit || *(volatile int*)P.status != 0) {
                        atomicCAS(P.status, 0, 104);
                        break;
                    }
                }
            }
            __threadfence();
        }
        if (threadIdx.x == 0) abort_flag = *(volatile int*)P.status != 0 ? 1 : 0;
        __syncthreads();
        if (abort_flag) return;
        const int r0 = K * 32;
        const int depth = __ldg(bdepth + K);
        if (P.epoch) { const long long c = clock64(); tw += c - c0; c0 = c; nlev_sum += depth + 1; }
        // step 1, all rows of the block at once (they do not depend on each other here): the near eliminations with rows
        // of EARLIER blocks -- those are final.  A row with local dependencies then stays in its group's buffer ...
        constexpr bool kPersist = (512 / G) >= 32;  // (one group per row of the block: the buffer survives until step 2)
        for (int g = grp; g < 32; g += ng) {
            const int i = r0 + g;
            if (i < P.n) {
                const int nfar = __ldg(nfar_of + i);
                const int nloc = (__ldg(nptr + i + 1) - __ldg(nptr + i)) - __ldg(crec + i);
                const int nl = P.dpos[i] - P.ptr[i];
                const bool small = P.ptr[i + 1] - P.ptr[i] <= RowSmem<G>::L;
                const bool keep = kPersist && small && nloc > 0;
                if (nl - nloc > nfar || nloc == 0 || keep)
                    factor_row_range<G, EXACT>(P, i, gl, gmask, R, big, nfar, nl - nloc, nfar > 0, true, !keep, keep ? nl : -1);
            }
        }
        __syncthreads();
        if (P.epoch) { const long long c = clock64(); tstep1 += c - c0; c0 = c; }
        // ... step 2, level by level of the block's local DAG: the eliminations with rows of the block itself
        for (int lv = 1; lv <= depth; ++lv) {
            bool hasbig = false;
            for (int g = grp; g < 32; g += ng) {
                const int i = r0 + g;
                if (i < P.n && __ldg(llev + i) == lv) {
                    const int nfar = __ldg(nfar_of + i);
                    const int nloc = (__ldg(nptr + i + 1) - __ldg(nptr + i)) - __ldg(crec + i);
                    const int nl = P.dpos[i] - P.ptr[i];
                    const bool small = P.ptr[i + 1] - P.ptr[i] <= RowSmem<G>::L;
                    const bool keep = kPersist && small;
                    hasbig = !small;
                    // (a row whose step 1 had nothing to do and was skipped still holds A's values)
                    const bool touched = nl - nloc > nfar || nfar > 0;
                    if (keep) ilu0_row_local<G, EXACT>(P, i, r0, gl, gmask, R, factor_smem, nl - nloc, nl);
                    else factor_row_range<G, EXACT>(P, i, gl, gmask, R, big, nl - nloc, nl, touched, true, true, -1);
                }
            }
            const int anybig = __syncthreads_or(hasbig ? 1 : 0);  // the rows of this local level are final (and visible) before the next level reads them
            if (P.epoch) { const long long c = clock64(); if (anybig) { tbig += c - c0; ++nbig; } else tl += c - c0; c0 = c; }
        }
        if (threadIdx.x == 0) {
            __threadfence();
            st_relaxed_gpu_u32(progress, (unsigned)(K + 1));
        }
    }
    if (P.epoch && threadIdx.x == 0) {
        printf("factor front: step 1 %.0f cycles per block\n", (double)tstep1 / nblocks);
        printf("factor front: %d blocks, %lld local levels | cycles per block: wait %.0f, levels without long rows %.0f, with %.0f | %lld levels with a long row, %.0f cycles each; %.0f cycles per other level\n",
               nblocks, nlev_sum, (double)tw / nblocks, (double)tl / nblocks, (double)tbig / nblocks, nbig, nbig ? (double)tbig / nbig : 0.0,
               (double)tl / (double)(nlev_sum - nbig));
    }
}

int factor_schedule(const ilub_handle* h) {
    const TriSolve& T = h->solve[0];
    const int algo = (int)options().factor_algo;
    // stencil matrices, fast arithmetic: the pivot recurrence on the line-systolic kernel (ilub_systolic.cu)
    if ((algo == 0 || algo == 4) && options().exact == 0 && T.sys.fbuilt) return 4;
    if (algo == 4) return 3;
    if (algo != 0) return algo;
    if (T.nlevels <= 4096 && (int64_t)h->n >= (int64_t)T.nlevels * 256) return 3;  // few wide levels: level loop
    if ((int64_t)T.nlevels * 64 > (int64_t)h->n) return 3;  // deep narrow DAG: row latency rules; the level loop and one
                                                             // launch per level measure alike (config 4: 2.9 / 3.2 s),
                                                             // sync-free is 3x slower there
    return 2;
}

template <int G>
static void launch_factor_block(bool exact, cudaStream_t s, const FactorParams& P, int nblocks, const int* llev, const int* bdepth,
                                const int* nfar, const int* nptr, const int* crec, unsigned* far_done, unsigned* progress, unsigned* next_block, unsigned epoch) {
    const int block = 512, stagers = 8;
    const size_t smem = (size_t)(block / G) * RowSmem<G>::kBytes + (P.big_cap > 0 ? (size_t)28 * P.big_cap + 16 : 0);
    if (exact) {
        ILUB_CUDA(cudaFuncSetAttribute(k_factor_block<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_factor_block<G, true><<<1 + stagers, block, smem, s>>>(P, nblocks, llev, bdepth, nfar, nptr, crec, far_done, progress, next_block, epoch);
    } else {
        ILUB_CUDA(cudaFuncSetAttribute(k_factor_block<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_factor_block<G, false><<<1 + stagers, block, smem, s>>>(P, nblocks, llev, bdepth, nfar, nptr, crec, far_done, progress, next_block, epoch);
    }
}

// block-sequential ILU(0): deep narrow DAGs whose L solve has the block layout of ilub_block.cu (its local levels)
bool factor_block(ilub_handle* h, cudaStream_t s) {
    const Options& opt = options();
    TriSolve& T = h->solve[0];
    if (opt.deep_block == 0 || h->kind != KIND_ILU0 || h->nsys != 1 || !T.blk.built || !T.blk.llev.p) return false;
    if (h->max_row_len > 4096) return false;  // (rows beyond the big row buffer keep the level loop's slow path)
    FactorParams P;
    memset(&P, 0, sizeof(P));
    P.n = h->n;
    P.ptr = h->ptr.p;
    P.idx = h->idx.p;
    P.dpos = h->dpos.p;
    P.a = h->a.p;
    P.lu = h->lu.p;
    P.status = h->status.p;
    P.spin_limit = (long long)opt.spin_limit_ms * h->clock_khz * 40;
    P.epoch = opt.trace != 0 ? 1u : 0u;  // (debug print of the front's time split)
    int G = (int)opt.factor_group;
    if (G != 8 && G != 16 && G != 32) {
        const double avg = (double)h->nnz / (double)h->n;
        G = avg <= 10 ? 8 : (avg <= 24 ? 16 : 32);
    }
    P.big_cap = h->max_row_len > 4 * G ? std::min(4096, (h->max_row_len + 63) / 64 * 64) : 0;
    const int nblocks = T.blk.nblocks;
    // hand-shake words: [0] progress, [1] next block of the stagers, [2 ..] far_done per block (epoch-tagged: never reset)
    if (h->fblock.n < (size_t)nblocks + 2) {
        h->fblock.alloc((size_t)nblocks + 2);
        ILUB_CUDA(cudaMemsetAsync(h->fblock.p, 0, sizeof(unsigned) * ((size_t)nblocks + 2), s));
        h->fblock_epoch = 0;
    }
    ILUB_CUDA(cudaMemsetAsync(h->fblock.p, 0, sizeof(unsigned) * 2, s));
    if (h->fnfar.n < (size_t)h->n) {
        h->fnfar.alloc((size_t)h->n);
        k_factor_count_far<<<(h->n + 255) / 256, 256, 0, s>>>(P, h->fnfar.p);
        h->launches++;
    }
    const unsigned epoch = ++h->fblock_epoch;
    const bool exact = opt.exact != 0;
    switch (G) {
        case 8: launch_factor_block<8>(exact, s, P, nblocks, T.blk.llev.p, T.blk.bdepth.p, h->fnfar.p, T.blk.nptr.p, T.blk.cr.p, h->fblock.p + 2, h->fblock.p, h->fblock.p + 1, epoch); break;
        case 16: launch_factor_block<16>(exact, s, P, nblocks, T.blk.llev.p, T.blk.bdepth.p, h->fnfar.p, T.blk.nptr.p, T.blk.cr.p, h->fblock.p + 2, h->fblock.p, h->fblock.p + 1, epoch); break;
        default: launch_factor_block<32>(exact, s, P, nblocks, T.blk.llev.p, T.blk.bdepth.p, h->fnfar.p, T.blk.nptr.p, T.blk.cr.p, h->fblock.p + 2, h->fblock.p, h->fblock.p + 1, epoch); break;
    }
    ILUB_CUDA(cudaGetLastError());
    h->launches++;
    return true;
}

// symbolic work the chosen schedule needs on top of the level numbers: the (level, row) listing, or the chunked
// (part, level, row) order that the sync-free kernel walks
void factor_prepare(ilub_handle* h, cudaStream_t s) {
    if (h->max_row_len < 0) {
        DevBuf<int> m;
        m.alloc(1);
        ILUB_CUDA(cudaMemsetAsync(m.p, 0, sizeof(int), s));
        const int n1 = h->n / h->nsys;  // (the systems of a batch share one pattern)
        k_max_row_len<<<(n1 + 255) / 256, 256, 0, s>>>(n1, h->ptr.p, m.p);
        int v = 0;
        ILUB_CUDA(cudaMemcpyAsync(&v, m.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        ILUB_CUDA(cudaStreamSynchronize(s));
        h->max_row_len = v;
        h->launches++;
    }
    if ((options().factor_algo == 0 || options().factor_algo == 4) && options().exact == 0) stencil_factor_possible(h, s);
    if (factor_schedule(h) == 4) return;
    if (factor_schedule(h) == 2)
        analyse_generic(h, 0, s);
    else
        build_level_sets(h, 0, s);
}

void factor_numeric(ilub_handle* h, cudaStream_t s) {
    const Options& opt = options();
    TriSolve& T = h->solve[0];  // the factorisation DAG equals the L-solve DAG
    ILUB_REQUIRE(T.built, ILUB_ERR_INVALID, "factor_numeric before analysis");
    const bool ichol = h->kind == KIND_ICHOL0;
    FactorParams P;
    P.n = h->n;
    P.ptr = h->ptr.p;
    P.idx = h->idx.p;
    P.dpos = h->dpos.p;
    P.a = h->a.p;
    P.lu = h->lu.p;
    P.part_chunk_ptr = T.part_chunk_ptr.p;
    P.rows = T.rows.p;
    P.status = h->status.p;
    // a row of the sync-free schedule may wait for its whole upstream chain: the factorisation watchdog allows 40x the
    // solve limit (one minute by default)
    P.spin_limit = (long long)opt.spin_limit_ms * h->clock_khz * 40;
    P.big_cap = 0;  // (set below, once the lanes per row are known)
    P.nprefetch = 0;
    int G = (int)opt.factor_group;
    if (G != 4 && G != 8 && G != 16 && G != 32) {
        const double avg = (double)h->nnz / (double)h->n;
        G = avg <= 10 ? 8 : (avg <= 24 ? 16 : 32);
    }
    P.big_cap = h->max_row_len > 4 * G ? std::min(4096, (h->max_row_len + 63) / 64 * 64) : 0;
    const bool exact = opt.exact != 0;
    // schedule: 1 = one launch per level, 2 = sync-free, 3 = level loop (one cooperative launch), 0 = level loop when
    // the levels are few and wide (stencils, block-diagonal batches), sync-free otherwise
    factor_prepare(h, s);
    const int algo = factor_schedule(h);
    if (algo == 4) {
        factor_stencil(h, s);
        return;
    }
    if (algo == 3 && (int64_t)T.nlevels * 64 > (int64_t)h->n && factor_block(h, s)) {
        // (deep narrow DAG with the block layout at hand: block-sequential schedule)
    } else if (algo == 3) {
        if (h->fbarrier.n < 1) h->fbarrier.alloc(1);
        ILUB_CUDA(cudaMemsetAsync(h->fbarrier.p, 0, sizeof(unsigned), s));
        P.fflag = nullptr;
        P.epoch = 0;
        P.lvl_rows = T.lvl_rows.p;
        P.lvl_begin = P.lvl_end = 0;
        if (ichol)
            launch_levels<true>(G, exact, s, P, T.lvl_ptr.p, T.nlevels, h->fbarrier.p, h->sm_count, h->n);
        else
            launch_levels<false>(G, exact, s, P, T.lvl_ptr.p, T.nlevels, h->fbarrier.p, h->sm_count, h->n);
        h->launches++;
    } else if (algo == 2) {
        if (h->fflag.n < (size_t)h->n) {
            h->fflag.alloc(h->n);
            ILUB_CUDA(cudaMemsetAsync(h->fflag.p, 0, h->n, s));
            h->fepoch = 0;
        }
        if (++h->fepoch > 255) {
            ILUB_CUDA(cudaMemsetAsync(h->fflag.p, 0, h->n, s));
            h->fepoch = 1;
        }
        P.fflag = h->fflag.p;
        P.epoch = h->fepoch;
        P.lvl_rows = nullptr;
        P.lvl_begin = P.lvl_end = 0;
        // persistent: one CTA per part, all co-resident (T.nparts <= SMs * ctas_per_sm)
        const int block = 512;
        if (ichol)
            launch<true, true>(G, exact, T.nparts, block, s, P);
        else
            launch<false, true>(G, exact, T.nparts, block, s, P);
        h->launches++;
    } else {
        std::vector<int> lptr(T.nlevels + 1);
        ILUB_CUDA(cudaMemcpyAsync(lptr.data(), T.lvl_ptr.p, sizeof(int) * (T.nlevels + 1), cudaMemcpyDeviceToHost, s));
        ILUB_CUDA(cudaStreamSynchronize(s));
        P.fflag = nullptr;
        P.epoch = 0;
        P.lvl_rows = T.lvl_rows.p;
        const int block = 256;
        const int ng = block / G;
        for (int l = 0; l < T.nlevels; ++l) {
            P.lvl_begin = lptr[l];
            P.lvl_end = lptr[l + 1];
            const int cnt = P.lvl_end - P.lvl_begin;
            if (cnt <= 0) continue;
            const int grid = (cnt + ng - 1) / ng;
            if (ichol)
                launch<true, false>(G, exact, grid, block, s, P);
            else
                launch<false, false>(G, exact, grid, block, s, P);
            h->launches++;
        }
    }
    ILUB_CUDA(cudaGetLastError());
}

}  // namespace ilub
