// Plane-per-warp triangular solves for 2-D grids and block-diagonal batches of them (BASELINE config 5: 4096
// independent 2-D Poisson systems x 16 right-hand sides) -- the multi-right-hand-side path behind apply() / `B @ X` (replaces the column loop over matrix_sparse::triangular_solve,
// sparse_implementation.h:4040-4087, ilupp/__init__.py:122-150 `_matmat`).
//
// Same (task, step, lane) coefficient stream as the line-systolic kernel (ilub_systolic.cu): one lane per grid line,
// lanes skewed by one step, the offset-1 chain in a register, the offset-nx dependency one shuffle.  What differs is
// how the work is spread.  Planes of a 2-D problem are independent, so there is nothing to hand over between warps:
//   * ONE warp owns a whole plane for one right-hand side.  Its line blocks (32 lines each) run in the same warp as
//     independent instruction streams, line block jb delayed by 32 steps, so the value lane 0 of block jb needs from
//     lane 31 of block jb - 1 is what that lane produced one warp step earlier: the shuffle simply wraps around.
//     No rings, no polling, no flags -- and NB-fold instruction-level parallelism in a warp that used to issue one
//     dependent chain.
//   * The consumer warps of a CTA solve the SAME plane for different right-hand sides and share one copy of the
//     coefficient stages (1-D bulk copies issued by a producer warp, full/empty mbarrier pipeline): the factor is
//     read from HBM once per CTA instead of once per right-hand side; CTAs of the same plane are adjacent in the
//     grid, so the remaining re-reads are L2 hits.
//   * Right-hand side in / solution out: every warp moves its own data with 8-byte cp.async copies in skewed
//     4-line x 8-column rectangles (line group q is shifted by -4q columns, so a rectangle is consumed within 11
//     steps of its arrival): the window is 32 slots per line instead of 64, 8.25 KB per (line block, rhs).
//   * A CTA walks several consecutive planes: a line block starts the next plane as soon as it has finished the
//     previous one, so the ramps of the line-block skew are paid once per CTA, not once per plane.
// Arithmetic and rounding are those of the line-systolic kernel (`exact` = the reference's operation order, bit for
// bit; otherwise coefficients pre-divided by the diagonal and FMAs).
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "ilub_common.cuh"
#include "ilub_ptx.cuh"

namespace ilub {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kSteps = kSysStageSteps;            // steps per coefficient stage
constexpr int kStageLB = kSteps * 4 * 32 * 8;     // bytes of one stage of one line block (8 KB)
constexpr int kWin = 32;                          // window slots per line
constexpr int kWinStride = kWin + 1;              // odd stride: a warp's column accesses are conflict-free (stride 39 would also free the
                                                  // rectangle accesses of their 2-way conflicts: measured, no gain)
constexpr int kWinBytes = 32 * kWinStride * 8;    // one (line block, rhs) window
constexpr int kSkew = 4;                          // warp blocks between the starts of consecutive line blocks (32 steps)
constexpr int kMaxConsumers = 7;                    // + the producer warp = 256 threads: the register file allows 255 per thread

struct PlaneParams {
    const double* coef;
    double* x;
    long long ldx;
    long long n;
    int nrhs;
    int nx, ny, nyb, nplanes, nsteps_pad;
    int backward, order_xyz, unit_diag;
    int wpc;     // consumer warps per CTA
    int nst;     // coefficient stages
    int groups;  // CTAs per plane (groups of wpc * R right-hand sides)
    int npc;     // consecutive planes per CTA
    int* status;
    unsigned spin_iters;
};

__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds_2f64(uint32_t a, double& v0, double& v1) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a));
}
// 8-byte asynchronous copy global -> shared; src_bytes = 0 writes zeros (padding rows and columns)
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, unsigned src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8_if(uint32_t dst, const void* src, unsigned pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 8;\n\t}" ::"r"(dst), "l"(src), "r"(pred)
                 : "memory");
}
__device__ __forceinline__ void stg_cs_if(double* dst, double v, unsigned pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.cs.f64 [%0], %1;\n\t}" ::"l"(dst), "d"(v), "r"(pred) : "memory");
}
__device__ __forceinline__ void sts_f64_if(uint32_t a, double v, unsigned pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(a), "d"(v), "r"(pred) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void give_up(int* status, int code) { atomicCAS(status, 0, code); }

#define PLANE_SPIN(spins, code)                                        \
    if ((++(spins) & 0xfffu) == 0) {                                   \
        if ((spins) >= P.spin_iters || *(volatile int*)P.status != 0) { \
            give_up(P.status, code);                                   \
            return;                                                    \
        }                                                              \
    }

// position of a line block in the sequence of (plane, 8-step block) pairs a CTA walks through: g counts blocks from
// the line block's own start (negative: not started), m = block inside the current plane, off = element offset of
// (that plane, column 8 m) from the CTA's first plane
struct BlockPos {
    int g, m;
    long long off;
    __device__ __forceinline__ void init(int g0, long long off0) { g = g0; m = 0; off = off0; }
    __device__ __forceinline__ void advance(int nbl, long long step, long long wrap) {
        if (++g > 0) {
            off += step;
            if (++m == nbl) {
                m = 0;
                off += wrap;
            }
        }
    }
};

template <bool EXACT, int NB, int R>
__global__ void __launch_bounds__((kMaxConsumers + 1) * 32, 1) k_trsv_plane(const __grid_constant__ PlaneParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x % P.groups;
    const int p0 = (blockIdx.x / P.groups) * P.npc;     // first plane of this CTA
    const int np = min(P.npc, P.nplanes - p0);          // ... and how many it solves, one after the other
    const int rhs_cta = g * P.wpc * R;
    const int nactive = min(P.wpc, (P.nrhs - rhs_cta + R - 1) / R);  // consumer warps with at least one right-hand side
    // shared memory: [nst][NB] coefficient stages | [wpc][NB][R] windows | full[nst], empty[nst]
    unsigned char* const win0 = smem + (size_t)P.nst * NB * kStageLB;
    uint64_t* const full = reinterpret_cast<uint64_t*>(win0 + (size_t)P.wpc * NB * R * kWinBytes);
    uint64_t* const empty = full + P.nst;
    if ((int)threadIdx.x < P.nst) {
        mbar_init(&full[threadIdx.x], 1);
        mbar_init(&empty[threadIdx.x], (unsigned)nactive);
    }
    fence_mbar_init();
    __syncthreads();

    const int nyb = P.nyb, nst = P.nst;
    const int nbl = P.nsteps_pad / kSteps;  // 8-step blocks of a line block in one plane
    // Line block jb runs kSkew blocks behind jb - 1, and it moves on to the next plane of the CTA as soon as it has
    // finished the previous one: apart from the ramps at the two ends every line block is busy in every warp block.
    const int TL = np * nbl;
    const int K = kSkew * (nyb - 1) + TL + 2;  // ... plus the blocks that only drain
    unsigned spins = 0;

    if (warp == P.wpc) {
        // ------------------------------------------------------------ producer: coefficient stages for everyone
        if (lane != 0) return;
        int use = 0, slot = 0;
        unsigned parity = 1;  // phase parity of the PREVIOUS use of the slot
        for (int k = 0; k < K; ++k) {
            int cnt = 0;
#pragma unroll
            for (int jb = 0; jb < NB; ++jb) {
                const int gl = k - kSkew * jb;
                if (jb < nyb && gl >= 0 && gl < TL) ++cnt;
            }
            if (cnt == 0) continue;
            if (use >= nst) {
                while (!mbar_try_wait(&empty[slot], parity)) {
                    __nanosleep(40);
                    PLANE_SPIN(spins, 221)
                }
            }
            mbar_arrive_expect_tx(&full[slot], (unsigned)(cnt * kStageLB));
#pragma unroll
            for (int jb = 0; jb < NB; ++jb) {
                const int gl = k - kSkew * jb;
                if (jb < nyb && gl >= 0 && gl < TL) {
                    const int pl = gl / nbl, kl = gl - pl * nbl;
                    bulk_g2s(smem + (size_t)(slot * NB + jb) * kStageLB,
                             P.coef + ((size_t)((p0 + pl) * nyb + jb) * P.nsteps_pad + (size_t)kl * kSteps) * 128, (unsigned)kStageLB,
                             &full[slot]);
                }
            }
            ++use;
            if (++slot == nst) {
                slot = 0;
                parity ^= 1u;
            }
        }
        return;
    }
    if (warp >= nactive) return;

    // ---------------------------------------------------------------- consumer: planes p0 ..., right-hand sides rhs0 ...
    const int rhs0 = rhs_cta + warp * R;
    const int nr = min(R, P.nrhs - rhs0);
    const int nx = P.nx, ny = P.ny;
    const long long ldx = P.ldx;
    const uint32_t win_s = smem_u32(win0 + (size_t)warp * NB * R * kWinBytes);
    const uint32_t win_lane = win_s + lane * (kWinStride * 8);
    const uint32_t stage_lane = smem_u32(smem) + lane * 16;
    // rectangle roles: lane = (line rl of a 4-line group, column rc of 8)
    const int rl = lane >> 3, rc = lane & 7;
    const long long dir = P.backward ? -1 : 1;
    const long long e0 = ((long long)p0 * ny + rl) * nx + rc;  // element (line rl, column rc) of plane p0, solve numbering
    double* const g0 = P.x + (P.backward ? P.n - 1 - e0 : e0) + (long long)rhs0 * ldx;
    const long long qstride = dir * (4 * (long long)nx - 4);  // next line group: 4 lines down, 4 columns back
    const long long step8 = dir * 8;                                        // next block of a plane: 8 columns on
    const long long wrap8 = dir * ((long long)ny * nx - 8 * (long long)nbl);  // ... and from its last block to the next plane
    const uint32_t rect_s = win_s + rl * (kWinStride * 8);
    // line groups of the (ragged) last line block that exist for this lane
    unsigned lm_last = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (32 * (nyb - 1) + 4 * q + rl < ny) lm_last |= 1u << q;
    // columns [0, nxe) of this lane's line exist (0: the line itself does not)
    unsigned nxe[NB];
#pragma unroll
    for (int jb = 0; jb < NB; ++jb) nxe[jb] = 32 * jb + lane < ny ? (unsigned)nx : 0u;

    // element (line block jb, group q, block m of a plane): line 32 jb + 4 q + rl, column 8 m - 4 q + rc, consumed at the
    // line block's step 8 g + rc + rl (any q; g = blocks since the line block started) -> window slot (that step) & 31.
    // Mask of the groups q whose element exists (0 <= column < nx, line < ny):
    auto rect_mask = [&](int jb, int m) -> unsigned {
        const int a = 8 * m + rc;  // column of group 0; group q sits 4 q columns to the left
        const int qhi = a >> 2, qlo = (a - nx + 4) >> 2;
        const unsigned hi = qhi >= 7 ? 0xffu : ((2u << qhi) - 1u);
        const unsigned lo = qlo <= 0 ? 0xffu : (qlo >= 8 ? 0u : (0xffu << qlo) & 0xffu);
        return hi & lo & (jb == nyb - 1 ? lm_last : 0xffu);
    };
    const int qs = (int)qstride;
    const bool full_last = (ny & 31) == 0;
    auto rect_issue = [&](int jb, int r, const BlockPos& bp) {
        const double* const b = g0 + bp.off + (long long)r * ldx;
        const uint32_t d = rect_s + (uint32_t)((jb * R + r) * kWinBytes) + (((8 * bp.g + rc + rl) & (kWin - 1)) << 3);
        if (8 * bp.m >= 28 && 8 * bp.m + 8 <= nx && (jb < nyb - 1 || full_last)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) cp_async8(d + q * (4 * kWinStride * 8), b + q * qs, 8u);
        } else {
            const unsigned mask = rect_mask(jb, bp.m);
#pragma unroll
            for (int q = 0; q < 8; ++q) cp_async8_if(d + q * (4 * kWinStride * 8), b + q * qs, mask & (1u << q));
        }
    };
    auto rect_drain = [&](int jb, int r, const BlockPos& bp) {
        double* const b = g0 + bp.off + (long long)r * ldx;
        const uint32_t d = rect_s + (uint32_t)((jb * R + r) * kWinBytes) + (((8 * bp.g + rc + rl) & (kWin - 1)) << 3);
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = lds_f64(d + q * (4 * kWinStride * 8));
        if (8 * bp.m >= 28 && 8 * bp.m + 8 <= nx && (jb < nyb - 1 || full_last)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) __stcs(b + q * qs, v[q]);
        } else {
            const unsigned mask = rect_mask(jb, bp.m);
#pragma unroll
            for (int q = 0; q < 8; ++q) stg_cs_if(b + q * qs, v[q], mask & (1u << q));
        }
    };

    double xprev[NB][R];
#pragma unroll
    for (int jb = 0; jb < NB; ++jb)
#pragma unroll
        for (int r = 0; r < R; ++r) xprev[jb][r] = 0.0;
    const int order_xyz = P.order_xyz, unit_diag = P.unit_diag;
    const int src_lane = (lane + 31) & 31;
    int slot = 0;
    unsigned phase = 0;
    BlockPos pin[NB], pcur[NB], pout[NB];  // rectangle copied in / block computed / rectangle written back
#pragma unroll
    for (int jb = 0; jb < NB; ++jb) {
        pin[jb].init(-kSkew * jb, dir * ((long long)jb * 32 * nx));
        pcur[jb].init(-2 - kSkew * jb, 0);
        pout[jb].init(-4 - kSkew * jb, dir * ((long long)jb * 32 * nx));
    }

    for (int k = -2; k < K; ++k) {
        __syncwarp();  // the previous block's solution values are in the window
        // ---- solution out: rectangles consumed two blocks ago
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            if (jb < nyb && pout[jb].g >= 0 && pout[jb].g < TL) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (R == 1 || r < nr) rect_drain(jb, r, pout[jb]);
            }
        }
        __syncwarp();  // ... before their slots are handed to the copies below
        // ---- right-hand side in: rectangles needed two blocks from now (they reuse the slots just drained)
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            if (jb < nyb && pin[jb].g >= 0 && pin[jb].g < TL) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (R == 1 || r < nr) rect_issue(jb, r, pin[jb]);
            }
        }
        cp_async_commit();
        cp_async_wait<2>();  // the group committed two iterations ago (this block's rectangles) has landed
        __syncwarp();

        // A line block outside its own range (not started yet / finished: only in the ramps at the two ends) runs along
        // on whatever its addresses hold and its results are dropped; steps of a lane that fall outside its line
        // (column < 0 or >= nx, line >= ny) produce 0.0, so nothing stale -- no NaN of a neighbouring plane -- ever
        // reaches an element that exists.  The step body has no branches: the chains of a step overlap.
        unsigned act[NB][R];
        bool any = false;
        uint32_t wa[NB][R];
        int col0[NB];
        unsigned nxb[NB];
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            const bool a = jb < nyb && pcur[jb].g >= 0 && pcur[jb].g < TL;
            any = any || a;
            col0[jb] = 8 * pcur[jb].m - lane;
            nxb[jb] = a ? nxe[jb] : 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                act[jb][r] = (a && (R == 1 || r < nr)) ? 1u : 0u;
                wa[jb][r] = win_lane + (jb * R + r) * kWinBytes + (uint32_t)((pcur[jb].g & 3) << 6);
            }
        }
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            pin[jb].advance(nbl, step8, wrap8);
            pcur[jb].advance(nbl, 0, 0);
            pout[jb].advance(nbl, step8, wrap8);
        }
        if (k < 0 || !any) continue;
        while (!mbar_try_wait(&full[slot], phase)) PLANE_SPIN(spins, 222)
        const uint32_t st = stage_lane + (uint32_t)(slot * NB) * kStageLB;
        // operands of a step are loaded one step ahead
        double cyn[NB], cxn[NB], dvn[NB], bbn[NB][R];
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            cyn[jb] = lds_f64(st + jb * kStageLB + 8);
            lds_2f64(st + jb * kStageLB + 512, cxn[jb], dvn[jb]);
#pragma unroll
            for (int r = 0; r < R; ++r) bbn[jb][r] = lds_f64(wa[jb][r]);
        }
#pragma unroll
        for (int u = 0; u < kSteps; ++u) {
            double rot[NB][R], cy_[NB], cx_[NB], dv[NB], bb[NB][R], xnew[NB][R];
#pragma unroll
            for (int jb = 0; jb < NB; ++jb) {
                cy_[jb] = cyn[jb];
                cx_[jb] = cxn[jb];
                dv[jb] = dvn[jb];
#pragma unroll
                for (int r = 0; r < R; ++r) bb[jb][r] = bbn[jb][r];
            }
            // what the neighbouring line produced one step ago; lane 0 takes lane 31 of the previous line block
#pragma unroll
            for (int jb = 0; jb < NB; ++jb)
#pragma unroll
                for (int r = 0; r < R; ++r) rot[jb][r] = __shfl_sync(FULL, xprev[jb][r], src_lane);
            if (u + 1 < kSteps) {
#pragma unroll
                for (int jb = 0; jb < NB; ++jb) {
                    cyn[jb] = lds_f64(st + jb * kStageLB + (u + 1) * 1024 + 8);
                    lds_2f64(st + jb * kStageLB + (u + 1) * 1024 + 512, cxn[jb], dvn[jb]);
#pragma unroll
                    for (int r = 0; r < R; ++r) bbn[jb][r] = lds_f64(wa[jb][r] + (u + 1) * 8);
                }
            }
#pragma unroll
            for (int jb = 0; jb < NB; ++jb) {
                const bool inside = (unsigned)(col0[jb] + u) < nxb[jb];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double xy = rot[jb][r];
                    if (lane == 0) xy = jb > 0 ? rot[jb > 0 ? jb - 1 : 0][r] : 0.0;
                    double xn;
                    if (EXACT) {
                        const bool ay = __double_as_longlong(cy_[jb]) == kSysAbsentBits;
                        const bool ax = __double_as_longlong(cx_[jb]) == kSysAbsentBits;
                        double acc = bb[jb][r];
                        if (order_xyz) {
                            if (!ax) acc = __dsub_rn(acc, __dmul_rn(cx_[jb], xprev[jb][r]));
                            if (!ay) acc = __dsub_rn(acc, __dmul_rn(cy_[jb], xy));
                        } else {
                            if (!ay) acc = __dsub_rn(acc, __dmul_rn(cy_[jb], xy));
                            if (!ax) acc = __dsub_rn(acc, __dmul_rn(cx_[jb], xprev[jb][r]));
                        }
                        xn = unit_diag ? acc : __ddiv_rn(acc, dv[jb]);
                    } else {
                        double acc = bb[jb][r] * dv[jb];
                        acc = fma(-cx_[jb], xprev[jb][r], acc);
                        xn = fma(-cy_[jb], xy, acc);
                    }
                    xnew[jb][r] = inside ? xn : 0.0;
                }
            }
#pragma unroll
            for (int jb = 0; jb < NB; ++jb)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    sts_f64_if(wa[jb][r] + u * 8, xnew[jb][r], act[jb][r]);
                    xprev[jb][r] = xnew[jb][r];
                }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (++slot == nst) {
            slot = 0;
            phase ^= 1u;
        }
    }
    cp_async_wait<0>();
}

typedef void (*plane_kern_t)(const PlaneParams);

template <bool EXACT, int R>
plane_kern_t plane_kernel_nb(int nb) {
    switch (nb) {
        case 1: return (plane_kern_t)k_trsv_plane<EXACT, 1, R>;
        case 2: return (plane_kern_t)k_trsv_plane<EXACT, 2, R>;
        case 4: return (plane_kern_t)k_trsv_plane<EXACT, 4, R>;
        default: return (plane_kern_t)k_trsv_plane<EXACT, 8, 1>;  // (8 line blocks leave shared memory for one right-hand side per warp)
    }
}
int nb_index(int nb) { return nb == 1 ? 0 : nb == 2 ? 1 : nb == 4 ? 2 : 3; }

plane_kern_t plane_kernel(bool exact, int nb, int r) {
    const plane_kern_t k = exact ? (r == 2 ? plane_kernel_nb<true, 2>(nb) : plane_kernel_nb<true, 1>(nb))
                                 : (r == 2 ? plane_kernel_nb<false, 2>(nb) : plane_kernel_nb<false, 1>(nb));
    static thread_local unsigned attr_set[16] = {0};
    const int ki = (exact ? 8 : 0) + (r == 2 ? 4 : 0) + nb_index(nb);
    if (first_use_on_this_device(attr_set, ki))
        ILUB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return k;
}

}  // namespace

// Plane-per-warp variant of run_systolic for grids without a plane-to-plane dependency; false: not applicable
// (3-D coupling, more than 8 line blocks per plane, or switched off), the caller keeps the line-systolic kernel.
bool run_plane(ilub_handle* h, int slot, double* x, int64_t ldx, int nrhs, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    SysSolve& Y = T.sys;
    const Options& opt = options();
    if (!Y.built || Y.zcoupled || Y.nyb > 8 || opt.sys_plane == 0 || opt.trace) return false;
    // One warp per (plane, right-hand side) walks a plane 2-4x slower than the line-systolic kernel's team of warps
    // (measured: 256^2, one right-hand side: 623 us against 155 us), so it pays when there is other work to overlap:
    // several right-hand sides (the line-systolic kernel takes them one launch after the other) or more planes than
    // two waves of CTAs
    if (opt.sys_plane == 1 && nrhs < 3 && Y.nplanes < 2 * h->sm_count) return false;
    const bool exact = opt.exact != 0;
    if (Y.packed_exact != (exact ? 1 : 0)) pack_systolic(h, slot, s);
    const int nb = Y.nyb <= 1 ? 1 : Y.nyb <= 2 ? 2 : Y.nyb <= 4 ? 4 : 8;
    int r = opt.plane_r == 2 && nb < 8 ? 2 : 1;
    const int nst = 2;
    const size_t budget = (size_t)227 * 1024 - 256;
    const size_t coef_bytes = (size_t)nst * nb * kStageLB;
    if (coef_bytes + (size_t)nb * r * kWinBytes + 64 > budget) r = 1;
    if (coef_bytes + (size_t)nb * r * kWinBytes + 64 > budget) return false;
    int wpc = (int)std::min<size_t>(kMaxConsumers, (budget - coef_bytes - 64) / ((size_t)nb * r * kWinBytes));
    if (opt.plane_wpc > 0) wpc = std::min<int>(wpc, (int)opt.plane_wpc);
    wpc = std::max(1, std::min(wpc, (nrhs + r - 1) / r));
    PlaneParams P;
    P.coef = Y.coef.p;
    P.x = x;
    P.ldx = ldx;
    P.n = h->n;
    P.nrhs = nrhs;
    P.nx = Y.nx;
    P.ny = Y.ny;
    P.nyb = Y.nyb;
    P.nplanes = Y.nplanes;
    P.nsteps_pad = Y.nsteps_pad;
    P.backward = T.backward ? 1 : 0;
    P.order_xyz = Y.order_xyz;
    P.unit_diag = T.unit_diag ? 1 : 0;
    P.wpc = wpc;
    P.nst = nst;
    P.groups = (nrhs + wpc * r - 1) / (wpc * r);
    {   // planes per CTA: a CTA that walks several planes pays the ramp of the line-block skew once; fewest estimated
        // warp blocks on the busiest SM (CTAs of this size: one per SM at a time when wpc is at its maximum)
        const int nbl = Y.nsteps_pad / kSteps;
        const int64_t per_sm = std::max<int64_t>(1, ((int64_t)budget + 1024) / (int64_t)(coef_bytes + (size_t)wpc * nb * r * kWinBytes + 64));
        int64_t best = INT64_MAX;
        int best_np = 1;
        for (int npc = 1; npc <= 64 && npc <= Y.nplanes; ++npc) {
            const int64_t ctas_ = (int64_t)((Y.nplanes + npc - 1) / npc) * P.groups;
            const int64_t waves = (ctas_ + h->sm_count * per_sm - 1) / (h->sm_count * per_sm);
            const int64_t cost = waves * ((int64_t)npc * nbl + kSkew * (Y.nyb - 1) + 2 + 3);  // + start-up latency of a CTA
            if (cost < best) {
                best = cost;
                best_np = npc;
            }
        }
        P.npc = opt.plane_npc > 0 ? (int)std::min<int64_t>(opt.plane_npc, Y.nplanes) : best_np;
    }
    P.status = h->status.p;
    P.spin_iters = (unsigned)std::min<long long>(0x7fffffffll, std::max<long long>(1 << 16, (long long)opt.spin_limit_ms * h->clock_khz / 64));
    const size_t smem = coef_bytes + (size_t)wpc * nb * r * kWinBytes + 2 * nst * sizeof(uint64_t) + 16;
    const plane_kern_t kern = plane_kernel(exact, nb, r);
    const int64_t ctas = (int64_t)((Y.nplanes + P.npc - 1) / P.npc) * P.groups;
    ILUB_REQUIRE(ctas < ((int64_t)1 << 31), ILUB_ERR_SIZE, "plane solve: too many CTAs");
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas, 1, 1);
    cfg.blockDim = dim3((wpc + 1) * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    ILUB_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
    h->launches++;
    ILUB_CUDA(cudaGetLastError());
    return true;
}

}  // namespace ilub
