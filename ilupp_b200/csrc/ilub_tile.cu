// Tile-systolic sparse triangular solves for plane-coupled 3-D grids -- the headline kernel behind apply() /
// apply_trans() (replaces matrix_sparse::triangular_solve, sparse_implementation.h:4040-4087) when the analysis
// (ilub_systolic.cu) finds a 7-point-stencil dependency structure: every dependency at offset 1, nx or nx*ny of a
// lexicographically numbered nx x ny x nplanes grid.
//
// Schedule.  A grid line (fixed j, k; nx rows chained by the offset-1 dependency) belongs to ONE lane, which walks it
// front to back with the previous value of the line in a register.  A warp owns a TILE of TY x TZ = 8 x 4 adjacent
// lines (lane = a + 8 b) and runs them skewed by a + b steps: at step s lane (a, b) resolves element i = s - a - b, so
// BOTH cross-line dependencies -- (i, j-1, k) from lane (a-1, b) and (i, j, k-1) from lane (a, b-1) -- were produced
// one step earlier inside the warp: two shuffles.  Only the 8 + 4 boundary lanes of a tile take a value from a
// neighbouring tile, through a shared-memory ring of 16-byte {x, step tag} entries that is polled ONE STEP AHEAD of
// its use, so the dependent chain of a step is shuffle -> select -> 2 FMAs and never contains a poll.  Against the
// line-systolic schedule (32 x 1 lines per warp, ilub_systolic.cu) the wavefront crosses (ny/8 + nz/4) tile
// boundaries instead of nz plane boundaries (46 instead of 127 hand-offs on the critical path of 128^3), the skew
// padding of the coefficient stream drops from 24 % to 8 %, and 12 instead of 33 values per warp step cross warps.
//
// Data movement is all TMA: the packed coefficients of a tile ([step][{z, y | x, d}][lane], 1 KB per step) arrive in
// 8 KB stages through 1-D bulk copies, the right-hand side comes in and the solution goes out in 8 x 8 x 4 boxes
// (cp.async.bulk.tensor.3d on a tensor map of the vector; out-of-range parts of ragged tiles are clipped by the
// hardware), all issued by lane 0 of the compute warp itself -- there are no mover warps.  Backward solves run on
// the mirrored numbering (same kernel, mirrored box coordinates and window addressing).
//
// Tiles of a CTA: wy x wz compute warps.  Hand-offs inside a CTA: the producer stores straight into the consumer's
// ring; flow control = the consumer's stage counter, read one stage ahead.  Hand-offs between CTAs: global mailboxes
// {x, {epoch, step + 1}} (written once per solve, never reset) that fetcher warps of the consuming CTA poll several
// entries ahead and forward into the same rings.
//
// Arithmetic: `exact` = separately rounded multiply / subtract in the reference's accumulation order and a true
// division (bit-for-bit with a non-FMA build of the reference); otherwise FMA on coefficients pre-divided by the
// diagonal.  FACTOR = the pivot recurrence of ILU(0) / IChol(0) on stencil matrices (factor_stencil, ilub_systolic.cu).
#include <cooperative_groups.h>
#include <cuda.h>

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "ilub_common.cuh"
#include "ilub_ptx.cuh"

namespace ilub {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr long long kAbsentBits = kSysAbsentBits;
constexpr int TY = 8, TZ = 4;            // lines of a tile along y and z
constexpr int DY = TY - 1, DZ = TZ - 1;  // a boundary value is produced DY (DZ) steps after the step its consumer wants it for
constexpr int kSteps = kSysStageSteps;   // steps per stage
constexpr int kBody = 2 * kSteps;         // steps per iteration of the compute loop (two stages)
#ifndef ILUB_TILE_GATE
#define ILUB_TILE_GATE 4
#endif
#ifndef ILUB_TILE_TRACE
#define ILUB_TILE_TRACE 0
#endif
constexpr int kGate = ILUB_TILE_GATE;    // steps per gate on the ring inputs = steps per forwarder group (divides kSteps)
constexpr int kPub = kGate;              // the compute warp publishes its step counter this often
constexpr int kCoefBytes = kSteps * 4 * 32 * 8;  // coefficients of one stage (8 KB)
constexpr int kBlockBytes = kSteps * 32 * 8;     // window block: 8 columns of the 32 lines (2 KB)
constexpr int kMaxW = 6;                         // compute warps per CTA
constexpr int kMaxFetch = 4;
constexpr int kFetchDepth = 8;                   // mailbox entries a fetcher lane polls per round trip

struct TileParams {
    const double* coef;
    int nx, ny, nplanes, ntz1;  // ntz1: tiles along z per system (batches: nplanes = nz1 * nsys, nz1 whole tiles)
    int nxp, xoff;              // nx rounded up to whole window blocks; first real column in solve order (backward: nxp - nx)
    int nty, ntz, nyp, nzp;     // tile grid; lines along y / planes along z rounded up to whole tiles
    int wy, wz, gy, gz;      // tiles per CTA, CTA grid (cy = blockIdx.y, cz = blockIdx.x; the launch grid is padded to whole clusters)
    int ccy, ccz;            // thread-block cluster: ccy x ccz CTAs; hand-offs inside a cluster go through distributed shared memory
    int nstages, nsteps_pad, nblk;  // stages of 8 steps; real window blocks = ceil(nx / 8)
    int nst, nwin, ring;     // coefficient stages / window blocks / ring entries per line, per warp
    int rstride;             // bytes between the lines of a ring (ring * 16 + 16: the lines start in different banks)
    int per_warp;            // shared-memory bytes per compute warp
    int nfetch;
    int order_xyz, unit_diag;
    ulonglong2* ymail;
    ulonglong2* zmail;
    unsigned long long epoch;
    int pdl;
    int* status;
    unsigned spin_iters;
    long long* trace;
    unsigned nap_fwd, nap_fetch, nap_tma;  // nanoseconds an idle helper warp sleeps between polls
    int dbg;  // debug (option tile_debug): 2 no TMA stores, 4 no hand-off stores, 8 no proxy fence, 16 no window loads -- wrong results, timing only
};

struct TileCtx {
    unsigned char* smem;
    uint64_t* mbar;  // [W][nst]
    int* steps;      // [kMaxW] steps completed by every compute warp, published every kGate steps; a multiple of 8 also
                     //         says that the window stores of that stage are fenced for the TMA engine
    int* always;     // INT_MAX: the "consumer" behind a mailbox never blocks
    int* fprog;      // [kMaxW] producer steps the forwarder of every tile has forwarded (INT_MAX: nothing to forward)
    unsigned char* dummy;  // ring of always-ready zero entries for lanes without a ring input
    int off_win, off_yring, off_zring;
    int cy, cz, W;
};
struct TileInfo {
    int A, B;
    bool exists, has_yin, has_zin, has_yout, has_zout;
};
__device__ __forceinline__ TileInfo tile_info(const TileParams& P, const TileCtx& C, int w) {
    TileInfo T;
    T.A = C.cy * P.wy + w % P.wy;
    T.B = C.cz * P.wz + w / P.wy;
    T.exists = w < C.W && T.A < P.nty && T.B < P.ntz;
    T.has_yin = T.A > 0;
    T.has_zin = (T.B % P.ntz1) != 0;
    T.has_yout = T.A + 1 < P.nty;
    T.has_zout = T.B + 1 < P.ntz && ((T.B + 1) % P.ntz1) != 0;
    return T;
}

__device__ __forceinline__ int ld_prog(const int* p) { return *(const volatile int*)p; }
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds_2f64(uint32_t a, double& v0, double& v1) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a));
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void ld_s128(uint32_t a, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.cta.shared.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}" : "=l"(lo), "=l"(hi) : "r"(a) : "memory");
}
// one 16-byte store {x, tag} through a GENERIC address: a ring in this CTA's shared memory or a mailbox in global memory
__device__ __forceinline__ void st_gen128(void* p, double x, unsigned long long tag) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.gpu.b128 [%0], q;\n\t}" ::"l"(p),
                 "l"(__double_as_longlong(x)), "l"(tag)
                 : "memory");
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.relaxed.cta.shared.b64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
// both hand-offs of a step from one 16-byte register image {x, tag}
__device__ __forceinline__ void st_handoffs(void* py, bool doy, void* pz, bool doz, double x, unsigned long long tag) {
    asm volatile(
        "{\n\t.reg .b128 q;\n\t.reg .pred p0, p1;\n\t"
        "setp.ne.u32 p0, %4, 0;\n\tsetp.ne.u32 p1, %5, 0;\n\t"
        "mov.b128 q, {%2, %3};\n\t"
        "@p0 st.relaxed.gpu.b128 [%0], q;\n\t"
        "@p1 st.relaxed.gpu.b128 [%1], q;\n\t}" ::"l"(py),
        "l"(pz), "l"(__double_as_longlong(x)), "l"(tag), "r"((unsigned)doy), "r"((unsigned)doz)
        : "memory");
}
// slot of producer step s0 + off (off < 16) of a ring whose stage-aligned bases are b0 (s0) and b1 (s0 + 8)
#define RING_SLOT(b0, b1, off) (((off) < kSteps ? (b0) : (b1)) + ((off) & 7) * 16)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"((unsigned long long)map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// A wait that has failed `limit` polls (or sees another warp's failure) gives up: ILUB_ERR_WATCHDOG, not a hang.
#define TILE_SPIN(spins, code)                                          \
    if ((++(spins) & 0xfffu) == 0) {                                    \
        if ((spins) >= P.spin_iters || *(volatile int*)P.status != 0) { \
            atomicCAS(P.status, 0, code);                               \
            return;                                                     \
        }                                                               \
    }

// ---------------------------------------------------------------- fetcher: global mailboxes -> rings of this CTA's
// first tile row (y inputs) and first tile column (z inputs).  One lane per incoming line; entries are indexed by the
// producer's step; every round polls the next kFetchDepth entries and forwards the valid prefix.
__device__ __forceinline__ void tile_fetch(const TileParams& P, const TileCtx& C, int f, int lane) {
    // (lines whose producer sits in another CTA of this cluster arrive through distributed shared memory)
    const int nyl = (C.cy > 0 && C.cy % P.ccy == 0) ? P.wz * TZ : 0;
    const int nzl = (C.cz > 0 && C.cz % P.ccz == 0) ? P.wy * TY : 0;
    const int ngroups = (nyl + nzl + 31) / 32;
    if (ngroups == 0) return;
    const int line = (f % ngroups) * 32 + lane;
    const ulonglong2* src = nullptr;
    uint32_t dst = 0;
    int cons = 0, lead = 0;
    if (line < nyl) {
        const int wz = line / TZ, b = line % TZ;
        const int A = C.cy * P.wy, B = C.cz * P.wz + wz;
        if (A < P.nty && B < P.ntz) {
            cons = wz * P.wy;
            src = P.ymail + ((size_t)(A - 1) * P.nzp + (B * TZ + b)) * P.nsteps_pad;
            dst = smem_u32(C.smem + (size_t)cons * P.per_warp + C.off_yring) + b * P.rstride;
            lead = DY;
        }
    } else if (line < nyl + nzl) {
        const int l = line - nyl;
        const int wy = l / TY, a = l % TY;
        const int A = C.cy * P.wy + wy, B = C.cz * P.wz;
        if (A < P.nty && B < P.ntz && (B % P.ntz1) != 0) {
            cons = wy;
            src = P.zmail + ((size_t)(B - 1) * P.nyp + (A * TY + a)) * P.nsteps_pad;
            dst = smem_u32(C.smem + (size_t)cons * P.per_warp + C.off_zring) + a * P.rstride;
            lead = DZ;
        }
    }
    const unsigned long long ehi = (P.epoch & 0xffffffffull) << 32;
    const unsigned ringmask = (unsigned)P.ring - 1u;
    int e = src ? 0 : P.nsteps_pad;  // first producer step this lane has not forwarded
    unsigned spins = 0;
    while (true) {
        // entry e overwrites entry e - ring, which the consumer reads in its step e - ring - lead
        const int limit = min(P.nsteps_pad, ld_prog(&C.steps[cons]) + P.ring + lead);
        unsigned long long xb[kFetchDepth], tg[kFetchDepth];
#pragma unroll
        for (int u = 0; u < kFetchDepth; ++u) {
            tg[u] = 0ull;
            xb[u] = 0ull;
            if (e + u < limit) ld_b128_gpu(src + e + u, xb[u], tg[u]);
        }
        int adv = 0;
#pragma unroll
        for (int u = 0; u < kFetchDepth; ++u) {
            if (adv == u && e + u < limit && tg[u] == (ehi | (unsigned long long)(e + u + 1))) {
                st_ring_s(dst + (((unsigned)(e + u) & ringmask) << 4), __longlong_as_double((long long)xb[u]), (unsigned long long)(e + u + 1));
                adv = u + 1;
            }
        }
        e += adv;
        if (__all_sync(FULL, e >= P.nsteps_pad)) break;
        if (!__any_sync(FULL, adv > 0)) __nanosleep(P.nap_fetch);  // helper warps must not crowd the compute warps out of the issue slots
        TILE_SPIN(spins, 302)
    }
}

// ---------------------------------------------------------------- forwarder: one warp per compute warp.  Lane
// (half, l) serves outgoing boundary line l of the tile (4 y lines, 8 z lines) for one half of every group of kGate
// steps: it follows the compute warp's step counter, reads the boundary values out of the producer's window and writes
// {x, {epoch, step + 1}} into the consumer's ring (same CTA) or into a global mailbox (other CTA) -- the compute warps
// themselves never execute a hand-off instruction.
template <bool BWD, bool FACTOR>
__device__ __forceinline__ void tile_forward(const TileParams& P, const TileCtx& C, int w, int lane) {
    constexpr int NL = TY + TZ, HALF = kGate / 2;
    const TileInfo T = tile_info(P, C, w);
    if (!T.exists || !(T.has_yout || T.has_zout) || (P.dbg & 4)) return;  // (fprog preset to INT_MAX)
    const int half = lane / NL, l = lane % NL;
    const bool isy = l < TZ;
    const bool live = lane < 2 * NL && (isy ? T.has_yout : T.has_zout);
    const int a = isy ? TY - 1 : l - TZ, b = isy ? l : TZ - 1;
    const int sigma = a + b, pl = a + TY * b;
    const int wy = w % P.wy, wz = w / P.wy;
    const uint32_t win_row = smem_u32(C.smem + (size_t)w * P.per_warp + C.off_win) + (BWD ? 31 - pl : pl) * 64;
    unsigned char* dst0 = nullptr;
    unsigned mask = (unsigned)P.ring - 1u;
    // Where the hand-offs go: the ring of the consuming warp -- in this CTA, or in the neighbouring CTA of the cluster
    // through distributed shared memory (same generic store) -- or a global mailbox when the consumer sits in another
    // cluster.  ycons / zcons: step counter of the reader of a RING we write (flow control; mailboxes never block).
    namespace cg = cooperative_groups;
    const int* ycons = C.always;
    const int* zcons = C.always;
    const int rank = (C.cz % P.ccz) + (C.cy % P.ccy) * P.ccz;
    const bool y_local = wy + 1 < P.wy, y_dsm = !y_local && (C.cy + 1) % P.ccy != 0;
    const bool z_local = wz + 1 < P.wz, z_dsm = !z_local && (C.cz + 1) % P.ccz != 0;
    const int ycw = y_local ? w + 1 : wz * P.wy;       // consuming warp of the y hand-offs (in its CTA)
    const int zcw = z_local ? w + P.wy : wy;           // ... of the z hand-offs
    if (T.has_yout) {
        if (y_local) ycons = &C.steps[ycw];
        else if (y_dsm) ycons = (const int*)cg::this_cluster().map_shared_rank((void*)&C.steps[ycw], rank + P.ccz);
    }
    if (T.has_zout) {
        if (z_local) zcons = &C.steps[zcw];
        else if (z_dsm) zcons = (const int*)cg::this_cluster().map_shared_rank((void*)&C.steps[zcw], rank + 1);
    }
    if (live) {
        if (isy) {
            unsigned char* ring = C.smem + (size_t)ycw * P.per_warp + C.off_yring + b * P.rstride;
            if (y_local) {
                dst0 = ring;
            } else if (y_dsm) {
                dst0 = (unsigned char*)cg::this_cluster().map_shared_rank((void*)ring, rank + P.ccz);
            } else {
                dst0 = (unsigned char*)(P.ymail + ((size_t)T.A * P.nzp + (T.B * TZ + b)) * P.nsteps_pad);
                mask = 0x0fffffffu;
            }
        } else {
            unsigned char* ring = C.smem + (size_t)zcw * P.per_warp + C.off_zring + a * P.rstride;
            if (z_local) {
                dst0 = ring;
            } else if (z_dsm) {
                dst0 = (unsigned char*)cg::this_cluster().map_shared_rank((void*)ring, rank + 1);
            } else {
                dst0 = (unsigned char*)(P.zmail + ((size_t)T.B * P.nyp + (T.A * TY + a)) * P.nsteps_pad);
                mask = 0x0fffffffu;
            }
        }
    }
    const unsigned long long ehi = (P.epoch & 0xffffffffull) << 32;
    const int nwin = P.nwin;
    unsigned spins = 0;
    int s = 0;                                   // first producer step of the group being forwarded
    int ip = -sigma + half * HALF;               // this lane's first column of the group (solve order; backward: padding first)
    int slot = (4 * nwin + (ip >> 3)) % nwin;    // window slot of block ip >> 3
    int yseen = 0, zseen = 0;  // last progress read from the ring readers (a remote read is a round trip through the cluster)
    while (s < P.nsteps_pad) {
        // the group's values are in the window, and the ring slots it overwrites have been read
        if (yseen < s + kGate - P.ring - DY) yseen = ld_prog(ycons);
        if (zseen < s + kGate - P.ring - DZ) zseen = ld_prog(zcons);
        if (ld_prog(&C.steps[w]) >= s + kGate && yseen >= s + kGate - P.ring - DY && zseen >= s + kGate - P.ring - DZ) {
            if (live) {
                const int slot_next = slot + 1 == nwin ? 0 : slot + 1;
                double v[HALF];
#pragma unroll
                for (int k = 0; k < HALF; ++k) {
                    const int ipk = ip + k, c = ipk & 7;
                    const int sk = ((ipk ^ ip) & ~7) == 0 ? slot : slot_next;
                    v[k] = 0.0;
                    if ((unsigned)(ipk - P.xoff) < (unsigned)P.nx) {
                        v[k] = lds_f64(win_row + sk * kBlockBytes + (BWD ? 7 - c : c) * 8);
                        if (FACTOR) v[k] = 1.0 / v[k];
                    }
                }
#pragma unroll
                for (int k = 0; k < HALF; ++k) {
                    const int e = s + half * HALF + k;
                    st_gen128(dst0 + (size_t)((unsigned)e & mask) * 16, v[k], ehi | (unsigned long long)(unsigned)(e + 1));
                }
                if ((((ip + kGate) ^ ip) & ~7) != 0) slot = slot_next;
                ip += kGate;
            }
            s += kGate;
            __syncwarp();
            if (lane == 0) *(volatile int*)&C.fprog[w] = s;
        } else {
            __nanosleep(P.nap_fwd);
            TILE_SPIN(spins, 303)
        }
    }
}

// ---------------------------------------------------------------- TMA warp: lane w moves the data of compute warp w --
// coefficient stages and right-hand-side blocks in (NST stages ahead), finished solution blocks out
template <bool BWD, bool FACTOR>
__device__ __forceinline__ void tile_tma(const CUtensorMap* tmap, const TileParams& P, const TileCtx& C, int lane) {
    const int w = lane;
    const TileInfo T = tile_info(P, C, w);
    const int nstages = P.nstages, nst = P.nst, nwin = P.nwin;
    unsigned char* mine = C.smem + (size_t)(T.exists ? w : 0) * P.per_warp;
    const uint32_t win_s = smem_u32(mine) + C.off_win;
    uint64_t* bar = C.mbar + (T.exists ? w : 0) * nst;
    const double* coef_task = P.coef + (size_t)(T.B * P.nty + T.A) * P.nsteps_pad * 128;
    // TMA box of window block m: columns [8m, 8m + 8) of the tile's lines.  Backward solves: mirrored, with the padding
    // of ragged extents first in solve order, so that no box starts at a negative coordinate
    const int c1 = BWD ? P.nyp - TY * (T.A + 1) : T.A * TY;
    const int c2 = BWD ? P.nzp - TZ * (T.B + 1) : T.B * TZ;
    auto issue_stage = [&](int qq) {  // coefficient stage qq and window block qq -> barrier of slot qq % nst
        const int sc = qq % nst;
        const bool blk = !FACTOR && qq < P.nblk && !(P.dbg & 16);
        mbar_arrive_expect_tx(&bar[sc], (unsigned)(kCoefBytes + (blk ? kBlockBytes : 0)));
        bulk_g2s(mine + (size_t)sc * kCoefBytes, coef_task + (size_t)qq * (kCoefBytes / 8), (unsigned)kCoefBytes, &bar[sc]);
        if (blk) tma_load_3d(win_s + (qq % nwin) * kBlockBytes, tmap, BWD ? P.nxp - kSteps * (qq + 1) : kSteps * qq, c1, c2, smem_u32(&bar[sc]));
    };
    auto store_block = [&](int m) {
        if (P.dbg & 2) return;
        tma_store_3d(tmap, BWD ? P.nxp - kSteps * (m + 1) : kSteps * m, c1, c2, win_s + (m % nwin) * kBlockBytes);
    };
    // the vector -- and, right after a (re)factorisation, the coefficient stream -- belong to the previous kernels of
    // the stream until those grids have completed: nothing of this solve reads global memory before this point
    if (P.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (T.exists)
        for (int qq = 0; qq < nst && qq < nstages; ++qq) issue_stage(qq);
    int q = T.exists ? 0 : nstages;  // next stage to retire
    bool pending = false;            // stage q retired, its refill (stage q + nst) not issued yet
    unsigned spins = 0;
    while (true) {
        if (q < nstages) {
            if (!pending && ld_prog(&C.steps[w]) >= kSteps * (q + 1)) {
                // block q - 2 of the window is complete (the last lane finishes column 8m + 7 in step 8m + 7 + 10)
                const int mb = q - 2;
                if (mb >= 0 && mb < P.nblk) store_block(mb);
                bulk_commit();
                if (q + nst < nstages) pending = true;
                else ++q;
            }
            if (pending) {
                // the refill lands in the window slot of block q - 3: the forwarders have read its boundary values
                // (producer steps < 8 q) and its store, issued one stage ago, has read the slot
                if (ld_prog(&C.fprog[w]) >= kSteps * q) {
                    bulk_wait_read<1>();
                    issue_stage(q + nst);
                    pending = false;
                    ++q;
                }
            }
        }
        if (__all_sync(FULL, q >= nstages)) break;
        __nanosleep(P.nap_tma);
        TILE_SPIN(spins, 308)
    }
    if (T.exists) {
        for (int mb = max(nstages - 2, 0); mb < P.nblk; ++mb) store_block(mb);
        bulk_commit();
        bulk_wait_all();
#if ILUB_TILE_TRACE
        if (P.trace) P.trace[(size_t)(T.B * P.nty + T.A) * 3 * (nstages + 1) + nstages] = global_ns();
#endif
    }
}

// ---------------------------------------------------------------- compute warp: one tile, one wavefront level per step.
// Nothing but the recurrence: operands from shared memory, two shuffles, the arithmetic, the window store.  A lone
// warp issues roughly one instruction every five cycles, so the instruction count of a step IS its duration: the loop
// body covers two coefficient stages (16 steps) to spread the per-iteration bookkeeping, every address is a base
// register plus an immediate, and everything that is not the recurrence lives in the helper warps.
template <bool EXACT, bool BWD, bool FACTOR>
__device__ __forceinline__ void tile_compute(const TileParams& P, const TileCtx& C, int warp, int lane) {
    const TileInfo T = tile_info(P, C, warp);
    if (!T.exists) return;
    const int a = lane % TY, b = lane / TY;
    const int sigma = a + b, sl = sigma & 7, sh = sigma >> 3;
    const int nstages = P.nstages, nst = P.nst, nwin = P.nwin;
    const unsigned ringmask = (unsigned)P.ring - 1u;
    unsigned char* mine = C.smem + (size_t)warp * P.per_warp;
    const uint32_t mine_s = smem_u32(mine);
    const uint32_t coef_s = mine_s + lane * 16;
    const uint32_t row_s = mine_s + C.off_win + (BWD ? 31 - lane : lane) * 64;
    const uint32_t dummy_s = smem_u32(C.dummy);
    const uint32_t yin_s = (a == 0 && T.has_yin) ? mine_s + C.off_yring + b * P.rstride : dummy_s;
    const uint32_t zin_s = (b == 0 && T.has_zin) ? mine_s + C.off_zring + a * P.rstride : dummy_s;
    uint64_t* bar = C.mbar + warp * nst;
    volatile int* mysteps = &C.steps[warp];
    const int order_xyz = P.order_xyz, unit_diag = P.unit_diag;
    unsigned spins = 0;
    double xprev = 0.0;
    int slot = 0;              // coefficient slot of the iteration's first stage
    unsigned phases = 0;       // bit k: parity of the next completion of slot k
    int sA = (4 * nwin - sh) % nwin;   // window slot of block q - sh (q = first stage of the iteration)
    int sB = (sA + nwin - 1) % nwin;   // ... of block q - sh - 1
#if ILUB_TILE_TRACE
    long long* trace = P.trace ? P.trace + (size_t)(T.B * P.nty + T.A) * 3 * (nstages + 1) : nullptr;
    long long wait_data = 0, wait_gate = 0;
#endif
    for (int q = 0; q < nstages; q += 2) {
#if ILUB_TILE_TRACE
        long long c0 = 0;
        if (trace) {
            if (lane == 0) {
                trace[q] = trace[q + 1] = global_ns();
                trace[(nstages + 1) + q] = trace[(nstages + 1) + q + 1] = wait_data;
                trace[2 * (nstages + 1) + q] = trace[2 * (nstages + 1) + q + 1] = wait_gate;
            }
            c0 = clock64();
        }
#endif
        const int slot2 = slot + 1 == nst ? 0 : slot + 1;
        // coefficient stage q and window block q have landed (stage q + 1 is waited for in the middle of the body)
        while (!mbar_try_wait(&bar[slot], (phases >> slot) & 1u)) TILE_SPIN(spins, 304)
#if ILUB_TILE_TRACE
        if (trace) wait_data += clock64() - c0;
#endif
        const int sC = sA + 1 == nwin ? 0 : sA + 1;
        const uint32_t st0 = coef_s + slot * kCoefBytes, st1 = coef_s + slot2 * kCoefBytes;
        // window address of step u: wB + 8u for u < sl, wA + 8u for sl <= u < sl + 8, wC + 8u beyond (blocks q - sh - 1,
        // q - sh, q - sh + 1); backward solves walk the mirrored block right to left
        const uint32_t wA = row_s + sA * kBlockBytes + (BWD ? (7 + sl) * 8 : -sl * 8);
        const uint32_t wB = row_s + sB * kBlockBytes + (BWD ? (sl - 1) * 8 : (8 - sl) * 8);
        const uint32_t wC = row_s + sC * kBlockBytes + (BWD ? (15 + sl) * 8 : -(sl + 8) * 8);
        const int s0 = q * kSteps;
        uint32_t yb[4], zb[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            yb[t] = yin_s + (((unsigned)(s0 + t * kSteps) & ringmask) << 4);
            zb[t] = zin_s + (((unsigned)(s0 + t * kSteps) & ringmask) << 4);
        }
        // iterations in which some lane is outside its line (i < 0 or i >= nx): window stores are predicated there only
        const bool edge = s0 < TY + TZ - 2 + P.xoff || s0 + kBody > P.nx + P.xoff;
        const int ibase = s0 - sigma - P.xoff;
        // Inputs from other tiles are gated in groups of kGate steps: the tag of the LAST entry a group needs is polled
        // (the entries of a line are written in step order by one thread, so the earlier ones are there as well), the
        // values themselves are plain loads.  In the last iteration the entries end at step nsteps_pad - 1: the gate
        // is clipped to it, and groups whose entries do not exist at all find their boundary lanes idle.
        const bool last = q + 2 >= nstages;
#define RING4(bs, off) ((bs)[(off) >> 3] + ((off) & 7) * 16)
        double cz_, cy_, cx_, dv, bb, ry_ = 0.0, rz_ = 0.0;
        unsigned long long gty = 0, gtz = 0;
        lds_2f64(st0, cz_, cy_);
        lds_2f64(st0 + 512, cx_, dv);
        bb = lds_f64(0 >= sl ? wA : wB);
        gty = lds_u64(RING4(yb, kGate - 1 + DY) + 8);
        gtz = lds_u64(RING4(zb, kGate - 1 + DZ) + 8);
#pragma unroll
        for (int u = 0; u < kBody; ++u) {
            if (u == kSteps - 1) {  // the second stage of the iteration (its first operands are requested during this step)
#if ILUB_TILE_TRACE
                if (trace) c0 = clock64();
#endif
                while (!mbar_try_wait(&bar[slot2], (phases >> slot2) & 1u)) TILE_SPIN(spins, 304)
#if ILUB_TILE_TRACE
                if (trace) wait_data += clock64() - c0;
#endif
            }
            if (u % kGate == 0) {
                const int gy = min(u + kGate - 1 + DY, last ? kBody - 1 : 99), gz = min(u + kGate - 1 + DZ, last ? kBody - 1 : 99);
                const int needy = (last && u + DY > kBody - 1) ? 0 : s0 + gy + 1, needz = (last && u + DZ > kBody - 1) ? 0 : s0 + gz + 1;
                if (!__all_sync(FULL, (int)(unsigned)gty >= needy && (int)(unsigned)gtz >= needz)) {
#if ILUB_TILE_TRACE
                    const long long g0 = trace ? clock64() : 0;
#endif
                    do {
                        TILE_SPIN(spins, 305)
                        gty = lds_u64((last ? RING4(yb, min(u + kGate - 1 + DY, kBody - 1)) : RING4(yb, u + kGate - 1 + DY)) + 8);
                        gtz = lds_u64((last ? RING4(zb, min(u + kGate - 1 + DZ, kBody - 1)) : RING4(zb, u + kGate - 1 + DZ)) + 8);
                    } while (!__all_sync(FULL, (int)(unsigned)gty >= needy && (int)(unsigned)gtz >= needz));
#if ILUB_TILE_TRACE
                    if (trace) wait_gate += clock64() - g0;
#endif
                }
                ry_ = lds_f64(RING4(yb, u + DY));
                rz_ = lds_f64(RING4(zb, u + DZ));
            }
            const double czu = cz_, cyu = cy_, cxu = cx_, dvu = dv, bbu = bb, ringy = ry_, ringz = rz_;
#define WIN_AT(v) (((v) < kSteps ? ((v) >= sl ? wA : wB) : ((v) - kSteps >= sl ? wC : wA)) + (BWD ? -(v) * 8 : (v) * 8))
            const uint32_t wu = WIN_AT(u);
            if (u + 1 < kBody) {  // next step's operands and inputs: independent of everything below
                const uint32_t stn = (u + 1 < kSteps ? st0 : st1) + ((u + 1) & 7) * 1024;
                lds_2f64(stn, cz_, cy_);
                lds_2f64(stn + 512, cx_, dv);
                bb = lds_f64(WIN_AT(u + 1));
                if ((u + 1) % kGate != 0) {
                    ry_ = lds_f64(RING4(yb, u + 1 + DY));
                    rz_ = lds_f64(RING4(zb, u + 1 + DZ));
                } else {  // the next group's gate, checked one step from now
                    gty = lds_u64((last ? RING4(yb, min(u + kGate + DY, kBody - 1)) : RING4(yb, u + kGate + DY)) + 8);
                    gtz = lds_u64((last ? RING4(zb, min(u + kGate + DZ, kBody - 1)) : RING4(zb, u + kGate + DZ)) + 8);
                }
            }
            double xy = __shfl_up_sync(FULL, xprev, 1);
            double xz = __shfl_up_sync(FULL, xprev, TY);
            if (a == 0) xy = ringy;
            if (b == 0) xz = ringz;
            double xnew;
            if (EXACT) {
                const bool az = __double_as_longlong(czu) == kAbsentBits;
                const bool ay = __double_as_longlong(cyu) == kAbsentBits;
                const bool ax = __double_as_longlong(cxu) == kAbsentBits;
                double acc = bbu;
                if (order_xyz) {
                    if (!ax) acc = __dsub_rn(acc, __dmul_rn(cxu, xprev));
                    if (!ay) acc = __dsub_rn(acc, __dmul_rn(cyu, xy));
                    if (!az) acc = __dsub_rn(acc, __dmul_rn(czu, xz));
                } else {
                    if (!az) acc = __dsub_rn(acc, __dmul_rn(czu, xz));
                    if (!ay) acc = __dsub_rn(acc, __dmul_rn(cyu, xy));
                    if (!ax) acc = __dsub_rn(acc, __dmul_rn(cxu, xprev));
                }
                xnew = unit_diag ? acc : __ddiv_rn(acc, dvu);
            } else {
                // coefficients are pre-divided by the diagonal (dvu = 1/d); absent entries carry 0.0.  FACTOR: the same
                // recurrence on the pivots -- d = a_ii - sum p_k / d_k with the products p_k in the coefficient slots
                // and a_ii in the diagonal slot; what travels between rows is 1 / d, what is stored is d
                double acc = FACTOR ? dvu : bbu * dvu;
                acc = fma(-cxu, xprev, acc);
                acc = fma(-cyu, xy, acc);
                xnew = fma(-czu, xz, acc);
            }
            if (!edge || (unsigned)(ibase + u) < (unsigned)P.nx) sts_f64(wu, xnew);
            if (FACTOR) xnew = 1.0 / xnew;
            xprev = xnew;
            if ((u + 1) % kSteps == 0) {
                // a stage is complete: its window stores become visible to the TMA engine (block q - 2 goes out)
                if (!(P.dbg & 8)) fence_proxy_async();
                __syncwarp();
                if (lane == 0) *mysteps = s0 + u + 1;
            } else if ((u + 1) % kPub == 0) {  // the forwarder may pick this group's boundary values up
                __syncwarp();
                if (lane == 0) *mysteps = s0 + u + 1;
            }
        }
#undef WIN_AT
#undef RING4
        phases ^= (1u << slot) | (1u << slot2);
        slot = slot2 + 1 == nst ? 0 : slot2 + 1;
        sB = sC;
        sA = sC + 1 == nwin ? 0 : sC + 1;
    }
#if ILUB_TILE_TRACE
    if (trace && lane == 0) {
        trace[(nstages + 1) + nstages] = wait_data;
        trace[2 * (nstages + 1) + nstages] = wait_gate;
    }
#endif
}

template <bool EXACT, bool BWD, bool FACTOR>
__global__ void __launch_bounds__(512, 1) k_trsv_tile(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TileCtx C;
    C.smem = smem;
    C.W = P.wy * P.wz;
    C.off_win = P.nst * kCoefBytes;
    C.off_yring = C.off_win + P.nwin * kBlockBytes;
    C.off_zring = C.off_yring + TZ * P.rstride;
    C.dummy = smem + (size_t)C.W * P.per_warp;
    C.mbar = reinterpret_cast<uint64_t*>(C.dummy + P.ring * 16);
    C.fprog = reinterpret_cast<int*>(C.mbar + kMaxW * 4);
    C.steps = C.fprog + kMaxW;
    C.always = C.steps + kMaxW;
    C.cy = blockIdx.y;
    C.cz = blockIdx.x;
    const int W = C.W;

    // windows zeroed (whatever an idle step reads is finite), rings zeroed (tag 0 = nothing yet), the dummy ring
    // always ready with value zero
    for (int w = 0; w < W; ++w) {
        ulonglong2* p = reinterpret_cast<ulonglong2*>(smem + (size_t)w * P.per_warp + C.off_win);
        const int cnt = (P.per_warp - C.off_win) / 16;
        for (int t = threadIdx.x; t < cnt; t += blockDim.x) p[t] = make_ulonglong2(0ull, 0ull);
    }
    for (int t = threadIdx.x; t < P.ring; t += blockDim.x) reinterpret_cast<ulonglong2*>(C.dummy)[t] = make_ulonglong2(0ull, (unsigned long long)INT_MAX);
    if (threadIdx.x < W * P.nst) mbar_init(&C.mbar[threadIdx.x], 1);
    if (threadIdx.x <= kMaxW) C.steps[threadIdx.x] = threadIdx.x == kMaxW ? INT_MAX : 0;
    if (threadIdx.x < kMaxW) {
        const TileInfo T = tile_info(P, C, threadIdx.x);
        C.fprog[threadIdx.x] = (T.exists && (T.has_yout || T.has_zout) && !(P.dbg & 4)) ? 0 : INT_MAX;
    }
    fence_mbar_init();
    fence_proxy_async();  // the zeroed windows are written again by the TMA engine
    __syncthreads();
    // every CTA of the cluster has initialised its rings before a neighbour writes into them
    if (P.ccy * P.ccz > 1) cooperative_groups::this_cluster().sync();
    // programmatic dependent launch: the next kernel of the stream (the second solve of an apply) may run its prologue
    // on idle SMs while this one runs; it touches global memory only behind griddepcontrol.wait
    if (P.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp < W) tile_compute<EXACT, BWD, FACTOR>(P, C, warp, lane);
    else if (warp < 2 * W) tile_forward<BWD, FACTOR>(P, C, warp - W, lane);
    else if (warp < 2 * W + P.nfetch) tile_fetch(P, C, warp - 2 * W, lane);
    else if (warp == 2 * W + P.nfetch) tile_tma<BWD, FACTOR>(&tmap, P, C, lane);
    // nobody leaves while a neighbour may still read its step counters or write into its rings
    if (P.ccy * P.ccz > 1) cooperative_groups::this_cluster().sync();
}

typedef void (*tile_kern_t)(const CUtensorMap, const TileParams);

tile_kern_t tile_kernel(bool exact, bool bwd, bool factor) {
    tile_kern_t k;
    int ki;
    if (factor) {
        k = (tile_kern_t)k_trsv_tile<false, false, true>;
        ki = 4;
    } else if (exact) {
        k = bwd ? (tile_kern_t)k_trsv_tile<true, true, false> : (tile_kern_t)k_trsv_tile<true, false, false>;
        ki = 2 + (bwd ? 1 : 0);
    } else {
        k = bwd ? (tile_kern_t)k_trsv_tile<false, true, false> : (tile_kern_t)k_trsv_tile<false, false, false>;
        ki = bwd ? 1 : 0;
    }
    static thread_local unsigned attr_set[5] = {0, 0, 0, 0, 0};
    if (first_use_on_this_device(attr_set, ki)) {
        ILUB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ILUB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    return k;
}

// how many clusters of ccy x ccz CTAs of this shape can be resident at once (0: not launchable)
int tile_active_clusters(tile_kern_t k, int ccy, int ccz, int gyp, int gzp, int block, size_t smem) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(gzp, gyp, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ccz;
    at[0].val.clusterDim.y = ccy;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

size_t tile_smem(int W, int nst, int nwin, int ring) {
    const size_t per_warp = (size_t)nst * kCoefBytes + (size_t)nwin * kBlockBytes + (size_t)(TY + TZ) * (ring * 16 + 16) + 64;
    return (size_t)W * per_warp + (size_t)ring * 16 + (size_t)kMaxW * 4 * 8 + (2 * kMaxW + 2) * 4 + 256;
}

typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_fn_t tensor_map_encoder() {
    static encode_fn_t fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        ILUB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        ILUB_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, ILUB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        fn = (encode_fn_t)p;
    }
    return fn;
}

}  // namespace

// CTA tiling of the tile grid; false: this grid does not fit (odd nx: the tensor map needs 16-byte strides; batches
// whose systems are not whole tiles deep; more tiles than one co-resident wave of compute warps)
bool tile_configure(ilub_handle* h, SysSolve& Y, int nsys) {
    const Options& opt = options();
    if (Y.nx % 2 != 0 || Y.nx < 2) return false;
    if (nsys > 1 && Y.nz1 % TZ != 0) return false;
    const int nty = (Y.ny + TY - 1) / TY;
    const int ntz1 = (Y.nz1 + TZ - 1) / TZ;
    const int64_t ntz = (int64_t)ntz1 * nsys;
    struct Cand {
        int wy, wz, nst, nwin, ring;
    };
    const Cand cands[] = {{2, 2, 3, 6, 64}, {2, 3, 2, 5, 32}, {3, 2, 2, 5, 32}};
    for (Cand c : cands) {
        if (opt.tile_nst > 0) {
            c.nst = (int)opt.tile_nst;
            c.nwin = c.nst + 3;
        }
        if (opt.tile_ring > 0) c.ring = (int)opt.tile_ring;
        if (opt.tile_warps > 0 && c.wy * c.wz != opt.tile_warps) continue;
        const int64_t gy = (nty + c.wy - 1) / c.wy, gz = (ntz + c.wz - 1) / c.wz;
        if (gy * gz > h->sm_count) continue;
        const size_t smem = tile_smem(c.wy * c.wz, c.nst, c.nwin, c.ring);
        if (smem > (size_t)227 * 1024) continue;
        Y.tile = 1;
        Y.ty = TY;
        Y.tz = TZ;
        Y.nty = nty;
        Y.ntz1 = ntz1;
        Y.ntz = (int)ntz;
        Y.twy = c.wy;
        Y.twz = c.wz;
        Y.tgy = (int)gy;
        Y.tgz = (int)gz;
        Y.tnst = c.nst;
        Y.tnwin = c.nwin;
        Y.tring = c.ring;
        Y.smem = smem;
        return true;
    }
    return false;
}

void launch_tile(ilub_handle* h, int slot, double* x, int64_t ldx, int nrhs, cudaStream_t s, const double* fcoef) {
    TriSolve& T = h->solve[slot];
    SysSolve& Y = T.sys;
    const Options& opt = options();
    const bool exact = fcoef ? false : opt.exact != 0;
    TileParams P;
    memset(&P, 0, sizeof(P));
    P.coef = fcoef ? fcoef : Y.coef.p;
    P.nx = Y.nx;
    P.ny = Y.ny;
    P.nplanes = Y.nplanes;
    P.ntz1 = Y.ntz1;
    P.nxp = (Y.nx + kSteps - 1) / kSteps * kSteps;
    P.xoff = (!fcoef && T.backward) ? P.nxp - Y.nx : 0;
    P.nty = Y.nty;
    P.ntz = Y.ntz;
    P.nyp = Y.nty * TY;
    P.nzp = Y.ntz * TZ;
    P.wy = Y.twy;
    P.wz = Y.twz;
    P.gy = Y.tgy;
    P.gz = Y.tgz;
    P.nsteps_pad = Y.nsteps_pad;
    P.nstages = Y.nsteps_pad / kSteps;
    P.nblk = (Y.nx + kSteps - 1) / kSteps;
    P.nst = Y.tnst;
    P.nwin = Y.tnwin;
    P.ring = Y.tring;
    P.rstride = Y.tring * 16 + 16;
    P.per_warp = Y.tnst * kCoefBytes + Y.tnwin * kBlockBytes + (TY + TZ) * P.rstride + 64;
    const int lines = (Y.tgy > 1 ? Y.twz * TZ : 0) + (Y.tgz > 1 ? Y.twy * TY : 0);
    const int groups = (lines + 31) / 32;
    P.nfetch = groups == 0 ? 0 : std::max(groups, (int)std::min<int64_t>(std::min(kMaxFetch, 15 - 2 * Y.twy * Y.twz), std::max<int64_t>(1, opt.tile_fetch)));
    P.order_xyz = Y.order_xyz;
    P.unit_diag = T.unit_diag ? 1 : 0;
    P.ymail = Y.ymail.p;
    P.zmail = Y.zmail.p;
    P.status = h->status.p;
    P.spin_iters = (unsigned)std::min<long long>(0x7fffffffll, std::max<long long>(1 << 16, (long long)opt.spin_limit_ms * h->clock_khz / 64));
    const bool bwd = fcoef ? false : T.backward;
    const tile_kern_t kern = tile_kernel(exact, bwd, fcoef != nullptr);
    const int block = (2 * Y.twy * Y.twz + P.nfetch + 1) * 32;
    if (Y.tccy == 0) {
        // thread-block cluster shape: the one with the fewest hand-offs through global memory on the critical path
        // among those whose (padded) grid is co-resident as whole clusters
        const int cmax = (int)std::min<int64_t>(16, opt.sys_cluster > 0 ? opt.sys_cluster : 16);
        long best = LONG_MAX;
        Y.tccy = Y.tccz = 1;
        for (int ccy = 1; ccy <= 4; ccy *= 2)
            for (int ccz = 1; ccz <= 8; ccz *= 2) {
                if (ccy * ccz > cmax || (ccy > 1 && ccy / 2 >= Y.tgy) || (ccz > 1 && ccz / 2 >= Y.tgz)) continue;
                const int gyp = (Y.tgy + ccy - 1) / ccy * ccy, gzp = (Y.tgz + ccz - 1) / ccz * ccz;
                if ((int64_t)gyp * gzp > h->sm_count) continue;
                const int nact = ccy * ccz > 1 ? tile_active_clusters(kern, ccy, ccz, gyp, gzp, block, Y.smem) : INT_MAX;
                if (getenv("ILUB_DEBUG")) fprintf(stderr, "[ilub] tile clusters %d x %d: %d clusters needed, %d can be active\n", ccy, ccz, gyp * gzp / (ccy * ccz), nact);
                if (gyp * gzp / (ccy * ccz) > nact) continue;
                const long cost = 1000L * ((gyp / ccy - 1) + (gzp / ccz - 1)) + (gyp * gzp - Y.tgy * Y.tgz);
                if (getenv("ILUB_DEBUG")) fprintf(stderr, "[ilub] tile clusters %d x %d: grid %d x %d, cost %ld\n", ccy, ccz, gyp, gzp, cost);
                if (cost < best) {
                    best = cost;
                    Y.tccy = ccy;
                    Y.tccz = ccz;
                }
            }
    }
    P.ccy = Y.tccy;
    P.ccz = Y.tccz;
    const int gyp = (Y.tgy + P.ccy - 1) / P.ccy * P.ccy, gzp = (Y.tgz + P.ccz - 1) / P.ccz * P.ccz;
    const size_t trace_len = (size_t)Y.ntasks * 3 * (P.nstages + 1);
    if (opt.trace && T.trace.n < trace_len) T.trace.alloc(trace_len);
    P.trace = opt.trace ? T.trace.p : nullptr;
    P.dbg = (int)opt.tile_debug;
    P.nap_fwd = (unsigned)opt.tile_nap_fwd;
    P.nap_fetch = (unsigned)opt.tile_nap_fetch;
    P.nap_tma = (unsigned)opt.tile_nap_tma;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(gzp, gyp, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = Y.smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = P.ccz;
    at[0].val.clusterDim.y = P.ccy;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    P.pdl = opt.sys_pdl != 0 ? 1 : 0;
    cfg.numAttrs = P.pdl ? 2 : 1;
    const encode_fn_t encode = tensor_map_encoder();
    for (int c = 0; c < nrhs; ++c) {
        double* xc = x + (int64_t)c * ldx;
        // the tensor map wants a 16-byte aligned vector: an odd-aligned column goes through the handle's workspace
        double* xt = xc;
        if (((uintptr_t)xc & 15u) != 0) {
            h->xalign.ensure((size_t)h->n);
            xt = h->xalign.p;
            ILUB_CUDA(cudaMemcpyAsync(xt, xc, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, s));
        }
        CUtensorMap tmap;
        const cuuint64_t dims[3] = {(cuuint64_t)Y.nx, (cuuint64_t)Y.ny, (cuuint64_t)Y.nplanes};
        const cuuint64_t strides[2] = {(cuuint64_t)Y.nx * 8, (cuuint64_t)Y.nx * Y.ny * 8};
        const cuuint32_t box[3] = {(cuuint32_t)kSteps, (cuuint32_t)TY, (cuuint32_t)TZ};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, xt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        ILUB_REQUIRE(r == CUDA_SUCCESS, ILUB_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
        P.epoch = ++Y.epoch;
        ILUB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, P));
        h->launches++;
        if (xt != xc) ILUB_CUDA(cudaMemcpyAsync(xc, xt, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, s));
    }
    ILUB_CUDA(cudaGetLastError());
}

}  // namespace ilub
