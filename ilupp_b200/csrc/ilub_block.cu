// Block-sequential sparse triangular solve for DEEP, NARROW, BAND-LOCAL dependency DAGs (BASELINE config 4: an irregular
// matrix whose solves have ~10^5 levels of ~9 rows, half of every row's dependencies within 64 rows of the diagonal) --
// replaces matrix_sparse::triangular_solve (sparse_implementation.h:4040-4087) for such matrices in the fast arithmetic
// mode; `exact` keeps the level loop of ilub_deep.cu, which reproduces the reference's accumulation order bit for bit.
//
// The level loop (ilub_deep.cu) pays one CTA barrier plus shared-memory round trips per LEVEL: 0.79 us x 113,837 levels.
// Here the rows are taken in their natural solve order in blocks of 32 consecutive rows, one row per lane of a compute
// warp (two warps take the blocks alternately: what a block needs from blocks older than the previous one is done while
// the previous one is still being finished), and nothing inside a block synchronises across warps:
//   * dependencies older than W = 512 rows ("far": the random long-range entries, 9.4 of the 19 per row) are final long
//     before the block is reached: stager warps, up to 11 blocks ahead of the compute warps, gather them from the solution vector
//     through L2 and reduce them into r_i = b_i / d_i - sum c_ij x_j, so they never touch the critical path;
//   * dependencies in the last W rows but in an earlier block ("recent") are a gather from a shared-memory ring of the
//     most recent 1024 solved values;
//   * dependencies inside the block ("local") form a tiny triangular system: x is held one value per lane and refined by
//     Jacobi sweeps x_i <- r_i - sum_local c_ij x_j with the x_j exchanged by warp shuffles; after s sweeps every row whose
//     local dependency chain has at most s links is exact, so depth(block) sweeps -- 4 on average for this matrix,
//     precomputed per block -- solve it exactly (same operations as forward substitution, no iteration error).
// The critical path is therefore ~n / 32 blocks of (ring gather + depth x (shuffle + FMA) rounds) instead of ~n / 9 levels
// of (barrier + shared-memory round trips).  Long rows (more near entries than a lane handles in registers) are rare and
// are reduced cooperatively by the whole warp.  Coefficients are pre-divided by the diagonal; the result differs from the
// reference's operation order by rounding only (tests hold it to 1e-12).
#include <algorithm>
#include <cstdio>

#include <cub/cub.cuh>

#include "ilub_common.cuh"
#include "ilub_ptx.cuh"

namespace ilub {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kBlk = 32;           // rows per block = lanes of the compute warp
constexpr int kWin = 512;          // rows of history a "recent" dependency may reach back (older ones are "far")
constexpr int kRing = 1024;        // shared-memory ring of solved values (>= kWin + 2 * kBlk, power of two)
#ifndef BLK_SLOTS
#define BLK_SLOTS 16
#endif
#ifndef BLK_STAGERS
#define BLK_STAGERS 8
#endif
constexpr int kSlots = BLK_SLOTS;         // staged blocks in flight: the stagers run up to kSlots - 1 blocks ahead
constexpr int kStagers = BLK_STAGERS;        // stager warps
static_assert(kSlots <= kWin / kBlk + 1, "a stager starts block K once block K - kSlots is done: its far dependencies (older than kWin rows) must be final by then");
constexpr int kSlotCap = 640;      // near entries of one block staged in shared memory (more: read from global)
constexpr int kRegLoc = 8;         // local entries a lane keeps in registers (more: the row is "long")
constexpr int kRegClo = 16;        // entries of a row's solved local system (closure) a lane keeps in registers
constexpr int kCloCap = 256;       // closure entries of one block staged in shared memory (more: the block runs the sweeps)
constexpr int kRegRec = 32;        // recent entries a lane gathers itself (more: the row is "long")

inline int blocks_for(int64_t work, int threads) { return (int)((work + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------ analysis
// solve position of a row / row of a solve position: q = backward ? n - 1 - row : row
__device__ __forceinline__ int mirror(int backward, int n, int v) { return backward ? n - 1 - v : v; }

// per solve position: how many far / recent / local entries
// ("recent" = older than the previous block, then the previous block itself: the two compute warps take the first part
//  one block ahead of time, see blk_compute)
__global__ void k_blk_count(int n, SrcDev S, int* __restrict__ cf, int* __restrict__ cr, int* __restrict__ cl,
                            int* __restrict__ co) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int row = mirror(S.backward, n, q);
    const int b0 = (q / kBlk) * kBlk;
    int f = 0, r = 0, l = 0, o = 0;
    for (int e = src_beg(S, row); e < src_end(S, row); ++e) {
        const int pd = mirror(S.backward, n, S.dep_idx[e]);
        if (pd >= b0) ++l;
        else if (pd >= b0 - kWin) {
            ++r;
            if (pd < b0 - kBlk) ++o;
        } else ++f;
    }
    cf[q] = f;
    cr[q] = r;
    cl[q] = l;
    co[q] = o;
}
__global__ void k_blk_sum(int n, const int* __restrict__ a, const int* __restrict__ b, int* __restrict__ c) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) c[q] = a[q] + b[q];
}
// entries in place: far list; near list = recent entries, then local entries
__global__ void k_blk_fill(int n, SrcDev S, const int* __restrict__ fptr, const int* __restrict__ nptr,
                           const int* __restrict__ cr, const int* __restrict__ co, int* __restrict__ fpos,
                           int* __restrict__ fsrc, int* __restrict__ npos, int* __restrict__ nsrc, int* __restrict__ dsrc) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int row = mirror(S.backward, n, q);
    const int b0 = (q / kBlk) * kBlk;
    int pf = fptr[q], po = nptr[q], pp = nptr[q] + co[q], pl = nptr[q] + cr[q];
    for (int e = src_beg(S, row); e < src_end(S, row); ++e) {
        const int pd = mirror(S.backward, n, S.dep_idx[e]);
        const int vp = S.val_pos ? S.val_pos[e] : e;
        if (pd >= b0) {
            npos[pl] = pd;
            nsrc[pl++] = vp;
        } else if (pd >= b0 - kBlk) {
            npos[pp] = pd;
            nsrc[pp++] = vp;
        } else if (pd >= b0 - kWin) {
            npos[po] = pd;
            nsrc[po++] = vp;
        } else {
            fpos[pf] = pd;
            fsrc[pf++] = vp;
        }
    }
    dsrc[q] = S.diag_pos ? S.diag_pos[row] : -1;
}
// per block: depth of the local dependency DAG (sweeps needed), widest recent / local list among the rows that keep that
// list in registers, and two masks of "long" rows: more than kRegRec recent entries (the warp gathers them cooperatively)
// and more than kRegLoc local entries (the lane walks its own list in shared memory).  One warp per block.
__global__ void k_blk_meta(int n, int nblocks, const int* __restrict__ nptr, const int* __restrict__ cr,
                           const int* __restrict__ co, const int* __restrict__ npos, int4* __restrict__ bmeta, int* __restrict__ llev,
                           int* __restrict__ bdepth, unsigned* __restrict__ cmask, int* __restrict__ ccnt) {
    const int K = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (K >= nblocks) return;
    const int q = K * kBlk + lane;
    const bool live = q < n;
    const int nb = live ? nptr[q] : 0, ne = live ? nptr[q + 1] : 0;
    const int nrec = live ? cr[q] : 0, nloc = ne - nb - nrec;
    const bool lrec = nrec > kRegRec, lloc = nloc > kRegLoc;
    const int widest = __reduce_max_sync(FULL, nloc);
    unsigned direct = 0;
    for (int t = 0; t < nloc; ++t) direct |= 1u << (npos[nb + nrec + t] - K * kBlk);
    // local levels, and the transitive closure of the local dependencies (the pattern of the block's solved local system),
    // by relaxation: at most 31 rounds
    int ll = 0;
    unsigned reach = direct;
    for (int it = 0; it < kBlk; ++it) {
        int nl = 0;
        unsigned nr = direct;
        for (int t = 0; t < widest; ++t) {  // (uniform trip count: the shuffles below need every lane)
            const bool has = t < nloc;
            const int src = has ? npos[nb + nrec + t] - K * kBlk : 0;
            const int v = __shfl_sync(FULL, ll, src);
            const unsigned m = __shfl_sync(FULL, reach, src);
            if (has) {
                nl = max(nl, v + 1);
                nr |= m;
            }
        }
        const bool changed = nl != ll || nr != reach;
        ll = nl;
        reach = nr;
        if (!__any_sync(FULL, changed)) break;
    }
    // a block whose closure would not fit its slot keeps the sweeps: nothing is stored for it (bit 30 of the first meta word;
    // this also bounds the closure arrays by kCloCap / kBlk entries per row)
    const bool noclo = __reduce_add_sync(FULL, live ? __popc(reach) : 0) > kCloCap;
    if (live) {
        cmask[q] = noclo ? 0u : reach;
        ccnt[q] = noclo ? 0 : __popc(reach);
    }
    const int depth_all = __reduce_max_sync(FULL, ll);
    const int depth_reg = __reduce_max_sync(FULL, lloc ? 0 : ll);
    // are the rows with long local lists leaves of the block's local DAG (no local entry of any row points at them)?  Then
    // their local sum is only needed once, after the last sweep
    const unsigned pointed = __reduce_or_sync(FULL, direct);
    const int nold = live ? co[q] : 0;
    const int maxold = __reduce_max_sync(FULL, lrec ? 0 : nold);
    const int maxprev = __reduce_max_sync(FULL, lrec ? 0 : nrec - nold);
    const int maxrec = maxold | (maxprev << 16);  // (widest older / previous-block list among the rows that gather their own)
    const int maxloc = __reduce_max_sync(FULL, lloc ? 0 : nloc);
    const unsigned mrec = __ballot_sync(FULL, lrec), mloc = __ballot_sync(FULL, lloc);
    const int leaf = (mloc & pointed) == 0 ? 1 : 0;
    const int depth = leaf ? depth_reg : depth_all;  // (leaves are evaluated behind the sweeps: they need none themselves)
    if (lane == 0) bmeta[K] = make_int4(depth | (leaf << 16) | (maxloc << 20) | ((noclo ? 1 : 0) << 30), maxrec, (int)mloc, (int)mrec);
    if (live) llev[q] = ll;
    if (lane == 0) bdepth[K] = depth_all;
}
// coefficient values in the layout's order, pre-divided by the row's diagonal
__global__ void k_blk_pack(int n, const int* __restrict__ fptr, const int* __restrict__ nptr, const int* __restrict__ fsrc,
                           const int* __restrict__ nsrc, const int* __restrict__ dsrc, const double* __restrict__ lu,
                           double* __restrict__ fval, double* __restrict__ nval, double* __restrict__ rdiag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = (int)(t >> 3), gl = (int)(t & 7);
    if (q >= n) return;
    const double dinv = dsrc[q] >= 0 ? 1.0 / lu[dsrc[q]] : 1.0;
    if (gl == 0) rdiag[q] = dinv;
    for (int e = fptr[q] + gl; e < fptr[q + 1]; e += 8) fval[e] = lu[fsrc[e]] * dinv;
    for (int e = nptr[q] + gl; e < nptr[q + 1]; e += 8) nval[e] = lu[nsrc[e]] * dinv;
}

// closure positions of a row = the set bits of its mask, ascending
__global__ void k_blk_clo_fill(int n, const int* __restrict__ cptr, const unsigned* __restrict__ cmask, int* __restrict__ cpos) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    unsigned m = cmask[q];
    int p = cptr[q];
    while (m) {
        cpos[p++] = __ffs(m) - 1;
        m &= m - 1;
    }
}
// values of the solved local systems: N = (I + L_local)^-1 - I, row by row in solve order (every local dependency of a row is
// an earlier row of the block): N_i = - sum_j l_ij (e_j + N_j).  One warp per block, lane = COLUMN of N (columns are
// independent, so the row loop needs no synchronisation); the dense 32 x 32 block lives in shared memory.
constexpr int kCloWarps = 4;
__global__ void __launch_bounds__(kCloWarps * 32) k_blk_clo_pack(int n, int nblocks, const int* __restrict__ nptr,
                                                                const int* __restrict__ cr, const int* __restrict__ npos,
                                                                const double* __restrict__ nval, const int* __restrict__ cptr,
                                                                const int* __restrict__ cpos, double* __restrict__ cval) {
    __shared__ double Ms[kCloWarps][kBlk * 33];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = blockIdx.x * kCloWarps + w;
    if (K >= nblocks) return;
    double* M = Ms[w];
    const int b0 = K * kBlk;
    const int q = b0 + lane;
    const bool live = q < n;
    const int my_lb = live ? nptr[q] + cr[q] : 0, my_le = live ? nptr[q + 1] : 0;
    const int my_cn = live ? cptr[q + 1] - cptr[q] : 0;
    if (__reduce_max_sync(FULL, my_cn) == 0) return;  // (no local dependency in the block, or a block that keeps the sweeps)
    for (int i = 0; i < kBlk; ++i) {
        const int lb = __shfl_sync(FULL, my_lb, i), le = __shfl_sync(FULL, my_le, i);
        double acc = 0.0;
        for (int e = lb; e < le; ++e) {
            const int j = __ldg(npos + e) - b0;
            acc = fma(-__ldg(nval + e), (lane == j ? 1.0 : 0.0) + M[j * 33 + lane], acc);
        }
        M[i * 33 + lane] = acc;
    }
    __syncwarp();
    if (live)
        for (int e = cptr[q]; e < cptr[q + 1]; ++e) cval[e] = M[lane * 33 + cpos[e]];
}

// ------------------------------------------------------------------------------------------------ kernel
struct BlockParams {
    int n, nblocks, backward;
    const int* fptr;
    const int* fpos;
    const double* fval;
    const int* nptr;
    const int* cr;  // recent entries per position (the rest of its near list is local)
    const int* co;  // ... of which older than the previous block (they come first)
    const int* npos;
    const double* nval;
    const double* rdiag;
    const int4* bmeta;
    const int* cptr;  // solved local systems (closure lists)
    const int* cpos;
    const double* cval;
    double* x;  // right-hand side in, solution out (natural row order); CTA c takes column c: x + c * ldx
    long long ldx;
    int* status;
    unsigned spin_iters;
    int trace;  // debug (option trace): the compute warp prints where its cycles went
};

struct BlockSlot {
    double r[kBlk];            // b_i / d_i minus the far dependencies
    double nval[kSlotCap];     // near entries of the block (all rows, concatenated)
    int npos[kSlotCap];
    int off[kBlk];             // first near entry of every row inside the slot (or inside the global arrays: big blocks)
    int nrec[kBlk], nloc[kBlk], nold[kBlk];
    int4 meta;                 // {depth | leaf << 16 | widest local list << 20 | no closure << 30, widest older | previous-block list, long-local mask, long-recent mask}
    int big;                   // the block's near entries did not fit: read them from global memory
    double cval[kCloCap];      // the block's solved local system (closure lists, all rows concatenated)
    int cpos[kCloCap];
    int coff[kBlk], ccnt[kBlk];
    int clo;                   // ... is staged (else: the block runs Jacobi sweeps over its local entries)
};
struct BlockSmem {
    double ring[kRing];
    double xs[2][kBlk];        // the block's current values, per compute warp (rows with long local lists read them here)
    BlockSlot slot[kSlots];
    int staged[kSlots];        // block number + 1 whose data the slot holds
    int progress;              // blocks completed by the compute warp
    long long tick[2];
};

#define BLK_SPIN(spins, code)                                            \
    if ((++(spins) & 0x3ffu) == 0) {                                     \
        if ((spins) >= P.spin_iters || *(volatile int*)P.status != 0) {   \
            atomicCAS(P.status, 0, code);                                \
            return;                                                      \
        }                                                                \
    }

// (the rare paths -- long rows -- are kept out of line: inlined into every variant of the block routine they are most of its
//  code, and a lone warp that runs every instruction once per block pays for instruction fetch)
#ifdef BLK_INLINE_RARE
#define BLK_RARE __device__ __forceinline__
#else
#define BLK_RARE __device__ __noinline__
#endif
#ifdef BLK_NOINLINE_HOT
#define BLK_HOT __device__ __noinline__
#else
#define BLK_HOT __device__ __forceinline__
#endif
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// ---------------------------------------------------------------- stager: far dependencies + near lists of block K
// A stager is a chain of dependent HBM reads (pointers -> entries -> solved values), about 0.8 us each.  Everything that does
// not depend on the solve's progress -- pointers, counts, the first far entries, the near and closure lists -- is requested
// BEFORE the stager waits for its slot, in two rounds of independent loads; only the solved values and the slot writes
// come behind the wait.
#ifndef BLK_STAGENEAR
#define BLK_STAGENEAR 4
#endif
constexpr int kStageNear = BLK_STAGENEAR;  // near / closure entries per lane held in registers across the wait (x 32 lanes)
constexpr int kStageClo = 2;
#ifndef BLK_FARREG
#define BLK_FARREG 16
#endif
#ifndef BLK_HUB
#define BLK_HUB 4
#endif
constexpr int kHubUnroll = BLK_HUB;
constexpr int kFarReg = BLK_FARREG;  // far entries per lane requested before the wait
__device__ __forceinline__ void blk_stager(const BlockParams& P, BlockSmem& M, int w, int lane) {
    unsigned spins = 0;
    for (int K = w; K < P.nblocks; K += kStagers) {
        BlockSlot& L = M.slot[K % kSlots];
        const int q = K * kBlk + lane;
        const bool live = q < P.n;
        const int row = live ? (P.backward ? P.n - 1 - q : q) : 0;
        const int q_last = min(P.n, K * kBlk + kBlk);
        // round 1: pointers and per-row scalars
        const int fb = live ? __ldg(P.fptr + q) : 0, fe = live ? __ldg(P.fptr + q + 1) : 0;
        const int nb = live ? __ldg(P.nptr + q) : 0, ne = live ? __ldg(P.nptr + q + 1) : 0;
        const int cb = live ? __ldg(P.cptr + q) : 0, ce = live ? __ldg(P.cptr + q + 1) : 0;
        const int ne_all = __ldg(P.nptr + q_last), ce_all = __ldg(P.cptr + q_last);
        const int nrec = live ? __ldg(P.cr + q) : 0, nold = live ? __ldg(P.co + q) : 0;
        const double rd = live ? __ldg(P.rdiag + q) : 0.0;
        const double bq = live ? P.x[row] : 0.0;  // (the right-hand side: the row's own value is not solved yet)
        int4 meta = make_int4(0, 0, 0, 0);
        if (lane == 0) meta = __ldg(P.bmeta + K);
        const int nb0 = __shfl_sync(FULL, nb, 0), cb0 = __shfl_sync(FULL, cb, 0);
        const int total = ne_all - nb0, ctotal = ce_all - cb0;
        const bool big = total > kSlotCap;
        const bool clo = !big && ((__shfl_sync(FULL, meta.x, 0) >> 30) & 1) == 0 && ctotal <= kCloCap;
        const bool flong = fe - fb > 64;
        // round 2: entries
        int fp[kFarReg];
        double fc[kFarReg];
#pragma unroll
        for (int u = 0; u < kFarReg; ++u) {
            const bool has = !flong && fb + u < fe;
            fp[u] = has ? __ldg(P.fpos + fb + u) : -1;
            fc[u] = has ? __ldg(P.fval + fb + u) : 0.0;
        }
        int npr[kStageNear], cpr[kStageClo];
        double nvr[kStageNear], cvr[kStageClo];
#pragma unroll
        for (int u = 0; u < kStageNear; ++u) {
            const int t = lane + 32 * u;
            const bool has = !big && t < total;
            npr[u] = has ? __ldg(P.npos + nb0 + t) : 0;
            nvr[u] = has ? __ldg(P.nval + nb0 + t) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kStageClo; ++u) {
            const int t = lane + 32 * u;
            const bool has = clo && t < ctotal;
            cpr[u] = has ? __ldg(P.cpos + cb0 + t) : 0;
            cvr[u] = has ? __ldg(P.cval + cb0 + t) : 0.0;
        }
        // the slot is free once block K - kSlots is done; then every far dependency (older than kWin rows) is final too
        while (*(volatile int*)&M.progress < K - kSlots + 1) {
            __nanosleep(100);
            BLK_SPIN(spins, 401)
        }
        fence_acq_rel_cta();
        // round 3: solved values of the far dependencies; the entries of a following round (rows with more than kFarReg far
        // entries) are requested together with the solved values of the round before, so that a round costs one trip to
        // memory instead of two
        double r = bq * rd;
        int gp[8];
        double gc[8];
        int e0 = fb + kFarReg;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool has = !flong && e0 + u < fe;
            gp[u] = has ? __ldg(P.fpos + e0 + u) : -1;
            gc[u] = has ? __ldg(P.fval + e0 + u) : 0.0;
        }
        {
            double v[kFarReg];
#pragma unroll
            for (int u = 0; u < kFarReg; ++u) v[u] = fp[u] >= 0 ? __ldcg(P.x + (P.backward ? P.n - 1 - fp[u] : fp[u])) : 0.0;
            double r1 = 0.0;
#pragma unroll
            for (int u = 0; u < kFarReg; u += 2) {
                r = fma(-fc[u], v[u], r);
                r1 = fma(-fc[u + 1], v[u + 1], r1);
            }
            r += r1;
        }
        while (__any_sync(FULL, !flong && e0 < fe)) {
            double gv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) gv[u] = gp[u] >= 0 ? __ldcg(P.x + (P.backward ? P.n - 1 - gp[u] : gp[u])) : 0.0;
            int hp[8];
            double hc[8];
            e0 += 8;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool has = !flong && e0 + u < fe;
                hp[u] = has ? __ldg(P.fpos + e0 + u) : -1;
                hc[u] = has ? __ldg(P.fval + e0 + u) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) r = fma(-gc[u], gv[u], r);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                gp[u] = hp[u];
                gc[u] = hc[u];
            }
        }
        // hub rows: the whole warp strides over the row's far list
        unsigned fm = __ballot_sync(FULL, flong);
        while (fm) {
            const int src = __ffs(fm) - 1;
            fm &= fm - 1;
            const int b = __shfl_sync(FULL, fb, src), e = __shfl_sync(FULL, fe, src);
            double part = 0.0;
            for (int t = b + lane; t < e; t += 32 * kHubUnroll) {  // (kHubUnroll strides requested together)
                int pd[kHubUnroll];
                double c[kHubUnroll], v[kHubUnroll];
#pragma unroll
                for (int u = 0; u < kHubUnroll; ++u) {
                    const bool has = t + 32 * u < e;
                    pd[u] = has ? __ldg(P.fpos + t + 32 * u) : -1;
                    c[u] = has ? __ldg(P.fval + t + 32 * u) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kHubUnroll; ++u) v[u] = pd[u] >= 0 ? __ldcg(P.x + (P.backward ? P.n - 1 - pd[u] : pd[u])) : 0.0;
#pragma unroll
                for (int u = 0; u < kHubUnroll; ++u) part = fma(c[u], v[u], part);
            }
            part = warp_sum(part);
            if (lane == src) r -= part;
        }
        // the slot
#pragma unroll
        for (int u = 0; u < kStageNear; ++u) {
            const int t = lane + 32 * u;
            if (!big && t < total) {
                L.npos[t] = npr[u];
                L.nval[t] = nvr[u];
            }
        }
        if (!big)
            for (int t = lane + 32 * kStageNear; t < total; t += 32) {
                L.npos[t] = __ldg(P.npos + nb0 + t);
                L.nval[t] = __ldg(P.nval + nb0 + t);
            }
#pragma unroll
        for (int u = 0; u < kStageClo; ++u) {
            const int t = lane + 32 * u;
            if (clo && t < ctotal) {
                L.cpos[t] = cpr[u];
                L.cval[t] = cvr[u];
            }
        }
        if (clo)
            for (int t = lane + 32 * kStageClo; t < ctotal; t += 32) {
                L.cpos[t] = __ldg(P.cpos + cb0 + t);
                L.cval[t] = __ldg(P.cval + cb0 + t);
            }
        L.coff[lane] = cb - cb0;
        L.ccnt[lane] = ce - cb;
        L.r[lane] = r;
        L.off[lane] = big ? nb : nb - nb0;
        L.nrec[lane] = nrec;
        L.nloc[lane] = ne - nb - nrec;
        L.nold[lane] = nold;
        if (lane == 0) {
            L.big = big ? 1 : 0;
            L.clo = clo ? 1 : 0;
            L.meta = meta;
        }
        __syncwarp();
        fence_acq_rel_cta();
        if (lane == 0) *(volatile int*)&M.staged[K % kSlots] = K + 1;
    }
}

// ---------------------------------------------------------------- compute warps: the blocks, one after the other
// TWO warps take the blocks alternately.  A block is two phases: (A) everything that only needs blocks up to K - 2 --
// slot data into registers, the recent entries OLDER than the previous block -- and (B) what needs block K - 1: the
// entries that point into the previous block, the local sweeps, the publication.  Phase A of block K + 1 (other warp) runs
// beside phase B of block K, so only B is on the critical path.
// BIG = the block's near entries live in global memory (they did not fit the slot); NL = local entries per lane held in
// registers (2 / 4 / 6 / 8: every lane runs NL shuffles per sweep, absent entries carry coefficient 0).  A lone warp is
// bound by the latency of whatever it issues in sequence, so every phase requests all its operands first and the loops
// have compile-time trip counts (no branch, no divergence check in front of the shuffles).
template <bool BIG>
BLK_HOT double blk_gather(const BlockSmem& M, const int* np, const double* nv, int off, int cnt, int maxcnt,
                                             bool islong, double r) {
    for (int t0 = 0; t0 < maxcnt; t0 += 8) {  // eight entries per round: all indices, then all ring values, then the arithmetic
        int pp[8];
        double cc[8], vv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool has = !islong && t0 + u < cnt;
            pp[u] = has ? np[off + t0 + u] : 0;
            cc[u] = has ? nv[off + t0 + u] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) vv[u] = M.ring[pp[u] & (kRing - 1)];
        double r1 = 0.0;
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
            r = fma(-cc[u], vv[u], r);
            r1 = fma(-cc[u + 1], vv[u + 1], r1);
        }
        r += r1;
    }
    return r;
}
// long rows: the warp strides over entries [first, first + cnt) of every long row's list
BLK_RARE double blk_gather_long(const BlockSmem& M, const int* np, const double* nv, unsigned lmask, int first,
                                                  int cnt, int lane, double r) {
    while (lmask) {
        const int src = __ffs(lmask) - 1;
        lmask &= lmask - 1;
        const int o = __shfl_sync(FULL, first, src), c = __shfl_sync(FULL, cnt, src);
        double part = 0.0;
        for (int t = lane; t < c; t += 32) part = fma(nv[o + t], M.ring[np[o + t] & (kRing - 1)], part);
        part = warp_sum(part);
        if (lane == src) r -= part;
    }
    return r;
}

// a row with a long local list: the lane walks it itself, values from the block's copy in shared memory (two chains)
BLK_RARE double blk_local_long(const int* np, const double* nv, const double* xs, int first, int cnt, int b0) {
    double a0 = 0.0, a1 = 0.0;
    for (int t = 0; t < cnt; t += 4) {
        int pp[4];
        double cc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool has = t + u < cnt;
            pp[u] = has ? np[first + t + u] - b0 : 0;
            cc[u] = has ? nv[first + t + u] : 0.0;
        }
        a0 = fma(cc[0], xs[pp[0]], a0);
        a1 = fma(cc[1], xs[pp[1]], a1);
        a0 = fma(cc[2], xs[pp[2]], a0);
        a1 = fma(cc[3], xs[pp[3]], a1);
    }
    return a0 + a1;
}

template <bool BIG, int NL>
__device__ __forceinline__ bool blk_block(const BlockParams& P, BlockSmem& M, BlockSlot& L, int K, int cw, int lane, int depthw,
                                          int maxrec, unsigned mloc, unsigned mrec, unsigned& spins) {
    const int depth = depthw & 0xffff;
    const bool leaf = ((depthw >> 16) & 1) != 0;  // long local lists belong to leaves of the local DAG: summed once, at the end
    const int maxold = maxrec & 0xffff, maxprev = maxrec >> 16;
    const int* np = BIG ? P.npos : L.npos;
    const double* nv = BIG ? P.nval : L.nval;
    const int off = L.off[lane];
    const int nrec = L.nrec[lane], nloc = L.nloc[lane], nold = L.nold[lane];
    const bool lrec = (mrec >> lane) & 1u, lloc = (mloc >> lane) & 1u;
    double r = L.r[lane];
    const int b0 = K * kBlk;
    // ---- phase A: local entries into registers, the recent entries older than the previous block
    int lsrc[NL];
    double lco[NL];
#pragma unroll
    for (int t = 0; t < NL; ++t) {
        const bool has = !lloc && t < nloc;
        lsrc[t] = has ? np[off + nrec + t] - b0 : lane;
        lco[t] = has ? nv[off + nrec + t] : 0.0;
    }
    r = blk_gather<BIG>(M, np, nv, off, nold, maxold, lrec, r);
    if (mrec) r = blk_gather_long(M, np, nv, mrec, off, nold, lane, r);
    // ---- phase B: block K - 1 is complete
    while (*(volatile int*)&M.progress < K) {
        if ((++spins & 0x3ffu) == 0 && (spins >= P.spin_iters || *(volatile int*)P.status != 0)) {
            atomicCAS(P.status, 0, 403);
            return false;
        }
    }
    fence_acq_rel_cta();
    r = blk_gather<BIG>(M, np, nv, off + nold, nrec - nold, maxprev, lrec, r);
    if (mrec) r = blk_gather_long(M, np, nv, mrec, off + nold, nrec - nold, lane, r);
    // local dependencies: Jacobi sweeps over the block, values exchanged by shuffles
    double* xs = M.xs[cw];
    double x = r;
    for (int s = 0; s < depth; ++s) {
        double acc = r, acc1 = 0.0;
        if (mloc && !leaf) {  // (rare) a long local list that others depend on: its lane reads the block from shared memory
            xs[lane] = x;
            __syncwarp();
            if (lloc) acc -= blk_local_long(np, nv, xs, off + nrec, nloc, b0);
            __syncwarp();
        }
        double v[NL];
#pragma unroll
        for (int t = 0; t < NL; ++t) v[t] = __shfl_sync(FULL, x, lsrc[t]);
#pragma unroll
        for (int t = 0; t < NL; t += 2) {
            acc = fma(-lco[t], v[t], acc);
            acc1 = fma(-lco[t + 1], v[t + 1], acc1);
        }
        x = acc + acc1;
    }
    if (mloc && leaf) {
        xs[lane] = x;
        __syncwarp();
        if (lloc) x = r - blk_local_long(np, nv, xs, off + nrec, nloc, b0);
        __syncwarp();
    }
    // ---- publish
    const int q = b0 + lane;
    M.ring[q & (kRing - 1)] = x;
    if (q < P.n) __stcg(P.x + (P.backward ? P.n - 1 - q : q), x);
    __syncwarp();
    fence_acq_rel_cta();
    if (lane == 0) *(volatile int*)&M.progress = K + 1;
    return true;
}

// a row whose solved local system has more than kRegClo entries: the lane walks its list itself
BLK_RARE double blk_clo_long(const int* cp, const double* cv, const double* xs, int first, int cnt) {
    double a0 = 0.0, a1 = 0.0;
    for (int t = 0; t < cnt; t += 4) {
        int pp[4];
        double cc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool has = t + u < cnt;
            pp[u] = has ? cp[first + t + u] : 0;
            cc[u] = has ? cv[first + t + u] : 0.0;
        }
        a0 = fma(cc[0], xs[pp[0]], a0);
        a1 = fma(cc[1], xs[pp[1]], a1);
        a0 = fma(cc[2], xs[pp[2]], a0);
        a1 = fma(cc[3], xs[pp[3]], a1);
    }
    return a0 + a1;
}
// The block with its local system solved ahead of time (k_blk_clo_pack): x = r + N r, ONE exchange of the 32 right-hand
// sides instead of depth(block) dependent sweeps.  NC = closure entries per lane held in registers (8 / 16).
#ifdef BLK_TIMERS  // dev builds: where the cycles of a block go (printed with option trace)
__device__ long long g_blk_tick[2][6];
#define BLK_TICK(i)                                                   \
    {                                                                 \
        const long long c_ = clock64();                               \
        if (lane == 0) g_blk_tick[cw][i] += c_ - M.tick[cw];          \
        M.tick[cw] = c_;                                              \
    }
#define BLK_TICK_START M.tick[cw] = clock64();
#else
#define BLK_TICK(i)
#define BLK_TICK_START
#endif
template <int NC>
__device__ __forceinline__ bool blk_block_clo(const BlockParams& P, BlockSmem& M, BlockSlot& L, int K, int cw, int lane, int maxrec,
                                              unsigned mrec, unsigned mclo, unsigned& spins) {
    const int maxold = maxrec & 0xffff, maxprev = maxrec >> 16;
    const int* np = L.npos;
    const double* nv = L.nval;
    const int off = L.off[lane];
    const int nrec = L.nrec[lane], nold = L.nold[lane];
    const int coff = L.coff[lane], ccnt = L.ccnt[lane];
    const bool lrec = (mrec >> lane) & 1u, lclo = (mclo >> lane) & 1u;
    double r = L.r[lane];
    const int b0 = K * kBlk;
    BLK_TICK_START
    // ---- phase A: the closure list into registers (a long list: its first NC entries), the recent entries older than the
    // previous block
    int csrc[NC];
    double cco[NC];
#pragma unroll
    for (int t = 0; t < NC; ++t) {
        const bool has = t < ccnt;
        csrc[t] = has ? L.cpos[coff + t] : lane;
        cco[t] = has ? L.cval[coff + t] : 0.0;
    }
    r = blk_gather<false>(M, np, nv, off, nold, maxold, lrec, r);
    if (mrec) r = blk_gather_long(M, np, nv, mrec, off, nold, lane, r);
    BLK_TICK(0)
    // ---- phase B: block K - 1 is complete
    while (*(volatile int*)&M.progress < K) {
        if ((++spins & 0x3ffu) == 0 && (spins >= P.spin_iters || *(volatile int*)P.status != 0)) {
            atomicCAS(P.status, 0, 403);
            return false;
        }
    }
    fence_acq_rel_cta();
    BLK_TICK(1)
    r = blk_gather<false>(M, np, nv, off + nold, nrec - nold, maxprev, lrec, r);
    if (mrec) r = blk_gather_long(M, np, nv, mrec, off + nold, nrec - nold, lane, r);
    BLK_TICK(2)
    double* xs = M.xs[cw];
    if (mclo) {
        xs[lane] = r;
        __syncwarp();
    }
    double v[NC];
#pragma unroll
    for (int t = 0; t < NC; ++t) v[t] = __shfl_sync(FULL, r, csrc[t]);
    double a[4] = {r, 0.0, 0.0, 0.0};
#pragma unroll
    for (int t = 0; t < NC; ++t) a[t & 3] = fma(cco[t], v[t], a[t & 3]);
    double x = (a[0] + a[1]) + (a[2] + a[3]);
    if (mclo) {  // (only NC = kRegClo gets here with a long list: the lane adds the entries beyond its registers)
        if (lclo) x += blk_clo_long(L.cpos, L.cval, xs, coff + kRegClo, ccnt - kRegClo);
        __syncwarp();
    }
    BLK_TICK(3)
    // ---- publish
    const int q = b0 + lane;
    M.ring[q & (kRing - 1)] = x;
    if (q < P.n) __stcg(P.x + (P.backward ? P.n - 1 - q : q), x);
    __syncwarp();
    fence_acq_rel_cta();
    if (lane == 0) *(volatile int*)&M.progress = K + 1;
    BLK_TICK(4)
    return true;
}

__device__ __forceinline__ void blk_compute(const BlockParams& P, BlockSmem& M, int cw, int lane) {
    unsigned spins = 0;
    long long t_wait = 0, t_block = 0;
    for (int K = cw; K < P.nblocks; K += 2) {
        BlockSlot& L = M.slot[K % kSlots];
        long long c0 = P.trace ? clock64() : 0;
        while (*(volatile int*)&M.staged[K % kSlots] != K + 1) BLK_SPIN(spins, 402)
        // phase A reads ring values of blocks up to K - 2
        while (*(volatile int*)&M.progress < K - 1) BLK_SPIN(spins, 404)
        fence_acq_rel_cta();  // (a volatile poll + one fence measured faster than ld.acquire polls: 125 against 185 cycles)
        if (P.trace) { const long long c = clock64(); t_wait += c - c0; c0 = c; }
        const int4 bm = L.meta;
        const int depth = bm.x, maxrec = bm.y, maxloc = (bm.x >> 20) & 0xf;
        const unsigned mloc = (unsigned)bm.z, mrec = (unsigned)bm.w;
        bool ok;
        if (L.clo) {
            const int ccnt = L.ccnt[lane];
            const unsigned mclo = __ballot_sync(FULL, ccnt > kRegClo);
            const int maxclo = __reduce_max_sync(FULL, min(ccnt, kRegClo));
            // (two variants only: with more, the instruction footprint of the compute warps costs more than the spared shuffles)
            if (maxclo <= 8) ok = blk_block_clo<8>(P, M, L, K, cw, lane, maxrec, mrec, mclo, spins);
            else
            ok = blk_block_clo<16>(P, M, L, K, cw, lane, maxrec, mrec, mclo, spins);
        } else if (L.big) ok = blk_block<true, 8>(P, M, L, K, cw, lane, depth, maxrec, mloc, mrec, spins);
        else if (maxloc <= 2) ok = blk_block<false, 2>(P, M, L, K, cw, lane, depth, maxrec, mloc, mrec, spins);
        else if (maxloc <= 4) ok = blk_block<false, 4>(P, M, L, K, cw, lane, depth, maxrec, mloc, mrec, spins);
        else if (maxloc <= 6) ok = blk_block<false, 6>(P, M, L, K, cw, lane, depth, maxrec, mloc, mrec, spins);
        else ok = blk_block<false, 8>(P, M, L, K, cw, lane, depth, maxrec, mloc, mrec, spins);
        if (!ok) return;
        if (P.trace) t_block += clock64() - c0;
    }
    if (P.trace && lane == 0) {
        const double nb = (double)((P.nblocks + 1 - cw) / 2);
        printf("block solve, compute warp %d: %d blocks | cycles per own block: wait for slot / older blocks %.0f, phases A + B (incl. the wait for the previous block) %.0f\n",
               cw, (int)nb, t_wait / nb, t_block / nb);
#ifdef BLK_TIMERS
        printf("  closure blocks, warp %d: phase A %.0f | wait for the previous block %.0f | entries into it %.0f | exchange %.0f | publish %.0f\n", cw,
               g_blk_tick[cw][0] / nb, g_blk_tick[cw][1] / nb, g_blk_tick[cw][2] / nb, g_blk_tick[cw][3] / nb, g_blk_tick[cw][4] / nb);
        for (int i = 0; i < 6; ++i) g_blk_tick[cw][i] = 0;
#endif
    }
}

constexpr int kWarps = 2 * kStagers;  // warps 0 and 1 compute; warps 4, 5, 8, 9 (their schedulers) stay idle; the other six stage
__global__ void __launch_bounds__(kWarps * 32, 1) k_trsv_block(const BlockParams P0) {
    // one CTA per right-hand side: the columns of a multi-vector apply are independent solves on different SMs
    BlockParams P = P0;
    P.x += (long long)blockIdx.x * P0.ldx;
    extern __shared__ __align__(16) unsigned char blk_smem_raw[];
    BlockSmem& M = *reinterpret_cast<BlockSmem*>(blk_smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < kRing; t += blockDim.x) M.ring[t] = 0.0;
    if (threadIdx.x < kSlots) M.staged[threadIdx.x] = 0;
    if (threadIdx.x == 0) M.progress = 0;
    __syncthreads();
    // the compute warps have their schedulers (warps 0, 4, 8 and 1, 5, 9) to themselves: every issue slot a stager takes
    // there is latency on the critical path
    if (warp < 2) blk_compute(P, M, warp, lane);
    else if ((warp & 3) >= 2) blk_stager(P, M, (warp >> 2) * 2 + ((warp & 3) - 2), lane);
}

}  // namespace
namespace {

// ------------------------------------------------------------------------------------------------ dependency levels
// The same schedule in the (max, +1) semiring: level(i) = max over the dependencies of level(dep) + 1.  The sync-free level
// propagation of ilub_analysis.cu costs 2.7 us per level on such a DAG (0.3 s per solve at n = 10^6); here the far maxima
// are reduced ahead of time by the stagers, recent ones come from a ring, and the block's local levels settle in
// depth(block) sweeps of 32-bit shuffles.
struct LevelParams {
    int n, nblocks, backward;
    const int* fptr;
    const int* fpos;
    const int* nptr;
    const int* cr;
    const int* npos;
    const int4* bmeta;
    int* level;  // [n] by original row
    int* nlev;   // max level + 1
    int* status;
    unsigned spin_iters;
};
struct LevelSlot {
    int m[kBlk];             // max over the far dependencies of (level + 1)
    int npos[kSlotCap];
    int off[kBlk], nrec[kBlk], nloc[kBlk];
    int4 meta;
    int big;
};
struct LevelSmem {
    int ring[kRing];
    int xs[kBlk];
    LevelSlot slot[kSlots];
    int staged[kSlots];
    int progress;
};
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
#define LEV_SPIN(spins, code)                                            \
    if ((++(spins) & 0x3ffu) == 0) {                                     \
        if ((spins) >= P.spin_iters || *(volatile int*)P.status != 0) {   \
            atomicCAS(P.status, 0, code);                                \
            return;                                                      \
        }                                                                \
    }

constexpr int kLevStagers = 9;  // (k_levels_block: one compute warp, nine stagers)
__device__ __forceinline__ void lev_stager(const LevelParams& P, LevelSmem& M, int w, int lane) {
    unsigned spins = 0;
    for (int K = w; K < P.nblocks; K += kLevStagers) {
        LevelSlot& L = M.slot[K % kSlots];
        while (*(volatile int*)&M.progress < K - kSlots + 1) {
            __nanosleep(100);
            LEV_SPIN(spins, 411)
        }
        fence_acq_rel_cta();
        const int q = K * kBlk + lane;
        const bool live = q < P.n;
        const int fb = live ? __ldg(P.fptr + q) : 0, fe = live ? __ldg(P.fptr + q + 1) : 0;
        int m = 0;
        const bool flong = fe - fb > 64;
        if (!flong) {
            for (int e0 = fb; e0 < fe; e0 += 8) {
                int v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    v[u] = -1;
                    if (e0 + u < fe) {
                        const int pd = __ldg(P.fpos + e0 + u);
                        v[u] = __ldcg(P.level + (P.backward ? P.n - 1 - pd : pd));
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) m = max(m, v[u] + 1);
            }
        }
        unsigned fm = __ballot_sync(FULL, flong);
        while (fm) {
            const int src = __ffs(fm) - 1;
            fm &= fm - 1;
            const int b = __shfl_sync(FULL, fb, src), e = __shfl_sync(FULL, fe, src);
            int part = 0;
            for (int t = b + lane; t < e; t += 32) {
                const int pd = __ldg(P.fpos + t);
                part = max(part, __ldcg(P.level + (P.backward ? P.n - 1 - pd : pd)) + 1);
            }
            part = warp_max(part);
            if (lane == src) m = max(m, part);
        }
        const int nb = live ? __ldg(P.nptr + q) : 0, ne = live ? __ldg(P.nptr + q + 1) : 0;
        const int nb0 = __shfl_sync(FULL, nb, 0);
        const int total = __ldg(P.nptr + min(P.n, K * kBlk + kBlk)) - nb0;
        const bool big = total > kSlotCap;
        if (!big)
            for (int t = lane; t < total; t += 32) L.npos[t] = __ldg(P.npos + nb0 + t);
        const int nrec = live ? __ldg(P.cr + q) : 0;
        L.m[lane] = m;
        L.off[lane] = big ? nb : nb - nb0;
        L.nrec[lane] = nrec;
        L.nloc[lane] = ne - nb - nrec;
        if (lane == 0) {
            L.big = big ? 1 : 0;
            L.meta = __ldg(P.bmeta + K);
        }
        __syncwarp();
        fence_acq_rel_cta();
        if (lane == 0) *(volatile int*)&M.staged[K % kSlots] = K + 1;
    }
}

__device__ __forceinline__ void lev_compute(const LevelParams& P, LevelSmem& M, int lane) {
    unsigned spins = 0;
    int top = 0;
    for (int K = 0; K < P.nblocks; ++K) {
        LevelSlot& L = M.slot[K % kSlots];
        while (*(volatile int*)&M.staged[K % kSlots] != K + 1) LEV_SPIN(spins, 412)
        fence_acq_rel_cta();
        const int4 bm = L.meta;
        // (the sweeps of the solve may skip leaves; a level needs every row: depth of ALL rows = regular depth, or one
        //  more when the long rows are leaves)
        const int depth = (bm.x & 0xffff) + ((bm.x >> 16) & 1);
        const int maxrec = (bm.y & 0xffff) + (bm.y >> 16);  // (bound of the recent loop: older + previous-block lists)
        const unsigned lmask = (unsigned)(bm.z | bm.w);  // (here a row long in either list is handled by the whole warp)
        const int* np = L.big ? P.npos : L.npos;
        const int off = L.off[lane], nrec = L.nrec[lane], nloc = L.nloc[lane];
        const bool islong = (lmask >> lane) & 1u;
        const int b0 = K * kBlk;
        int m = L.m[lane];
        for (int t0 = 0; t0 < maxrec; t0 += 8) {
            int pp[8], vv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) pp[u] = (!islong && t0 + u < nrec) ? np[off + t0 + u] : -1;
#pragma unroll
            for (int u = 0; u < 8; ++u) vv[u] = pp[u] >= 0 ? M.ring[pp[u] & (kRing - 1)] : -1;
#pragma unroll
            for (int u = 0; u < 8; ++u) m = max(m, vv[u] + 1);
        }
        unsigned lm = lmask;
        while (lm) {
            const int src = __ffs(lm) - 1;
            lm &= lm - 1;
            const int o = __shfl_sync(FULL, off, src), c = __shfl_sync(FULL, nrec, src);
            int part = 0;
            for (int t = lane; t < c; t += 32) part = max(part, M.ring[np[o + t] & (kRing - 1)] + 1);
            part = warp_max(part);
            if (lane == src) m = max(m, part);
        }
        int lsrc[kRegLoc];
        bool lhas[kRegLoc];
#pragma unroll
        for (int t = 0; t < kRegLoc; ++t) {
            lhas[t] = !islong && t < nloc;
            lsrc[t] = lhas[t] ? np[off + nrec + t] - b0 : lane;
        }
        int x = m;
        for (int s = 0; s < depth; ++s) {
            int acc = m;
            if (lmask) {
                M.xs[lane] = x;
                __syncwarp();
                unsigned l2 = lmask;
                while (l2) {
                    const int src = __ffs(l2) - 1;
                    l2 &= l2 - 1;
                    const int o = __shfl_sync(FULL, off + nrec, src), c = __shfl_sync(FULL, nloc, src);
                    int part = 0;
                    for (int t = lane; t < c; t += 32) part = max(part, M.xs[np[o + t] - b0] + 1);
                    part = warp_max(part);
                    if (lane == src) acc = max(acc, part);
                }
                __syncwarp();
            }
#pragma unroll
            for (int t = 0; t < kRegLoc; ++t) {
                const int v = __shfl_sync(FULL, x, lsrc[t]);
                if (lhas[t]) acc = max(acc, v + 1);
            }
            x = acc;
        }
        const int q = b0 + lane;
        M.ring[q & (kRing - 1)] = x;
        if (q < P.n) {
            __stcg(P.level + (P.backward ? P.n - 1 - q : q), x);
            top = max(top, x + 1);
        }
        __syncwarp();
        fence_acq_rel_cta();
        if (lane == 0) *(volatile int*)&M.progress = K + 1;
    }
    top = warp_max(top);
    if (lane == 0) *P.nlev = top;
}

constexpr int kLevWarps = 12;
__global__ void __launch_bounds__(kLevWarps * 32, 1) k_levels_block(const LevelParams P) {
    extern __shared__ __align__(16) unsigned char blk_smem_raw[];
    LevelSmem& M = *reinterpret_cast<LevelSmem*>(blk_smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < kRing; t += blockDim.x) M.ring[t] = 0;
    if (threadIdx.x < kSlots) M.staged[threadIdx.x] = 0;
    if (threadIdx.x == 0) M.progress = 0;
    __syncthreads();
    if (warp == 0) lev_compute(P, M, lane);
    else if ((warp & 3) != 0) lev_stager(P, M, warp - 1 - (warp >> 2), lane);
}

}  // namespace

// ================================================================================================== host
// dependency levels of solve `slot` through the block layout (which must exist); returns the number of levels
int levels_block(ilub_handle* h, int slot, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    BlockSolve& B = T.blk;
    const size_t smem = sizeof(LevelSmem);
    static thread_local unsigned attr_set[1] = {0};
    if (first_use_on_this_device(attr_set, 0))
        ILUB_CUDA(cudaFuncSetAttribute(k_levels_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DevBuf<int> nlev;
    nlev.alloc(1);
    LevelParams P;
    P.n = h->n;
    P.nblocks = B.nblocks;
    P.backward = T.backward ? 1 : 0;
    P.fptr = B.fptr.p;
    P.fpos = B.fpos.p;
    P.nptr = B.nptr.p;
    P.cr = B.cr.p;
    P.npos = B.npos.p;
    P.bmeta = B.bmeta.p;
    P.level = T.level.p;
    P.nlev = nlev.p;
    P.status = h->status.p;
    P.spin_iters = (unsigned)std::min<long long>(0x7fffffffll, std::max<long long>(1 << 20, (long long)options().spin_limit_ms * h->clock_khz / 32));
    k_levels_block<<<1, kLevWarps * 32, smem, s>>>(P);
    h->launches++;
    int v = 0;
    ILUB_CUDA(cudaMemcpyAsync(&v, nlev.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    return v;
}

bool block_wanted(const ilub_handle* h, int slot) {
    const Options& opt = options();
    if (opt.exact != 0 || opt.deep_block == 0) return false;
    return deep_wanted(h, slot) && h->nsys == 1;
}

void analyse_block(ilub_handle* h, int slot, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    BlockSolve& B = T.blk;
    if (B.built) return;
    const int n = h->n;
    const SolveSource src = solve_source(h, slot);
    const SrcDev S = to_dev(src);
    DevBuf<int> cf, cl;
    cf.alloc((size_t)n + 1);
    cl.alloc((size_t)n + 1);
    B.cr.alloc((size_t)n + 1);
    B.co.alloc((size_t)n + 1);
    DevBuf<int> cn;
    cn.alloc((size_t)n + 1);
    k_blk_count<<<blocks_for(n, 256), 256, 0, s>>>(n, S, cf.p, B.cr.p, cl.p, B.co.p);
    B.fptr.alloc((size_t)n + 1);
    B.nptr.alloc((size_t)n + 1);
    ILUB_CUDA(cudaMemsetAsync(cf.p + n, 0, sizeof(int), s));
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cf.p, B.fptr.p, n + 1, s);
    DevBuf<unsigned char> tmp;
    tmp.alloc(tb);
    ILUB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cf.p, B.fptr.p, n + 1, s));
    k_blk_sum<<<blocks_for(n, 256), 256, 0, s>>>(n, B.cr.p, cl.p, cn.p);
    ILUB_CUDA(cudaMemsetAsync(cn.p + n, 0, sizeof(int), s));
    size_t tb2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, cn.p, B.nptr.p, n + 1, s);
    if (tb2 > tmp.n) tmp.alloc(tb2);
    size_t tbb = tmp.n;
    ILUB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tbb, cn.p, B.nptr.p, n + 1, s));
    int ends[2] = {0, 0};
    ILUB_CUDA(cudaMemcpyAsync(&ends[0], B.fptr.p + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaMemcpyAsync(&ends[1], B.nptr.p + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    B.nfar = ends[0];
    B.nnear = ends[1];
    B.fpos.alloc((size_t)std::max(1, B.nfar));
    B.fsrc.alloc((size_t)std::max(1, B.nfar));
    B.fval.alloc((size_t)std::max(1, B.nfar));
    B.npos.alloc((size_t)std::max(1, B.nnear));
    B.nsrc.alloc((size_t)std::max(1, B.nnear));
    B.nval.alloc((size_t)std::max(1, B.nnear));
    B.dsrc.alloc((size_t)n);
    B.rdiag.alloc((size_t)n);
    k_blk_fill<<<blocks_for(n, 256), 256, 0, s>>>(n, S, B.fptr.p, B.nptr.p, B.cr.p, B.co.p, B.fpos.p, B.fsrc.p, B.npos.p, B.nsrc.p,
                                                 B.dsrc.p);
    B.nblocks = (n + kBlk - 1) / kBlk;
    B.bmeta.alloc((size_t)B.nblocks);
    B.llev.alloc((size_t)n);
    B.bdepth.alloc((size_t)B.nblocks);
    DevBuf<unsigned> cmask;
    cmask.alloc((size_t)n);
    k_blk_meta<<<blocks_for((int64_t)B.nblocks * 32, 256), 256, 0, s>>>(n, B.nblocks, B.nptr.p, B.cr.p, B.co.p, B.npos.p, B.bmeta.p,
                                                                       B.llev.p, B.bdepth.p, cmask.p, cn.p);
    // closure lists of the solved local systems
    B.cptr.alloc((size_t)n + 1);
    ILUB_CUDA(cudaMemsetAsync(cn.p + n, 0, sizeof(int), s));
    tbb = tmp.n;
    ILUB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tbb, cn.p, B.cptr.p, n + 1, s));
    ILUB_CUDA(cudaMemcpyAsync(&ends[0], B.cptr.p + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    B.nclo = ends[0];
    B.cpos.alloc((size_t)std::max(1, B.nclo));
    B.cval.alloc((size_t)std::max(1, B.nclo));
    k_blk_clo_fill<<<blocks_for(n, 256), 256, 0, s>>>(n, B.cptr.p, cmask.p, B.cpos.p);
    h->launches += 8;
    ILUB_CUDA(cudaStreamSynchronize(s));  // temporaries die here
    B.packed = false;
    B.built = true;
}

void pack_block(ilub_handle* h, int slot, cudaStream_t s) {
    BlockSolve& B = h->solve[slot].blk;
    if (!B.built) return;
    k_blk_pack<<<blocks_for((int64_t)h->n * 8, 256), 256, 0, s>>>(h->n, B.fptr.p, B.nptr.p, B.fsrc.p, B.nsrc.p, B.dsrc.p, h->lu.p,
                                                                 B.fval.p, B.nval.p, B.rdiag.p);
    k_blk_clo_pack<<<blocks_for(B.nblocks, kCloWarps), kCloWarps * 32, 0, s>>>(h->n, B.nblocks, B.nptr.p, B.cr.p, B.npos.p, B.nval.p,
                                                                              B.cptr.p, B.cpos.p, B.cval.p);
    h->launches += 2;
    B.packed = true;
}

void run_block(ilub_handle* h, int slot, double* x, int64_t ldx, int nrhs, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    BlockSolve& B = T.blk;
    if (!B.packed) pack_block(h, slot, s);
    const size_t smem = sizeof(BlockSmem);
    static thread_local unsigned attr_set[1] = {0};
    if (first_use_on_this_device(attr_set, 0))
        ILUB_CUDA(cudaFuncSetAttribute(k_trsv_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BlockParams P;
    P.n = h->n;
    P.nblocks = B.nblocks;
    P.backward = T.backward ? 1 : 0;
    P.fptr = B.fptr.p;
    P.fpos = B.fpos.p;
    P.fval = B.fval.p;
    P.nptr = B.nptr.p;
    P.cr = B.cr.p;
    P.co = B.co.p;
    P.npos = B.npos.p;
    P.nval = B.nval.p;
    P.rdiag = B.rdiag.p;
    P.bmeta = B.bmeta.p;
    P.cptr = B.cptr.p;
    P.cpos = B.cpos.p;
    P.cval = B.cval.p;
    P.status = h->status.p;
    P.spin_iters = (unsigned)std::min<long long>(0x7fffffffll, std::max<long long>(1 << 20, (long long)options().spin_limit_ms * h->clock_khz / 32));
    P.trace = options().trace != 0 ? 1 : 0;
    P.x = x;
    P.ldx = ldx;
    k_trsv_block<<<nrhs, kWarps * 32, smem, s>>>(P);
    h->launches++;
    ILUB_CUDA(cudaGetLastError());
}

}  // namespace ilub
