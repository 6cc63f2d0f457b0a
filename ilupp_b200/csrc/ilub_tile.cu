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
constexpr int kCoefBytes = kSteps * 4 * 32 * 8;  // coefficients of one stage (8 KB)
constexpr int kBlockBytes = kSteps * 32 * 8;     // window block: 8 columns of the 32 lines (2 KB)
constexpr int kMaxW = 6;                         // compute warps per CTA
constexpr int kMaxFetch = 4;

struct TileParams {
    const double* coef;
    int nx, ny, nplanes, ntz1;  // ntz1: tiles along z per system (batches: nplanes = nz1 * nsys, nz1 whole tiles)
    int nxp, xoff;              // nx rounded up to whole window blocks; first real column in solve order (backward: nxp - nx)
    int nty, ntz, nyp, nzp;     // tile grid; lines along y / planes along z rounded up to whole tiles
    int wy, wz, gy, gz;      // tiles per CTA, CTA grid (blockIdx.x = cy * gz + cz)
    int nstages, nsteps_pad, nblk;  // stages of 8 steps; real window blocks = ceil(nx / 8)
    int nst, nwin, ring;     // coefficient stages / window blocks / ring entries per line, per warp
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
};

struct TileCtx {
    unsigned char* smem;
    uint64_t* mbar;  // [W][nst]
    int* prog;       // [W] completed stages of every compute warp
    int* always;     // INT_MAX: the "consumer" behind a mailbox never blocks
    unsigned char* dummy;  // ring of always-ready zero entries for lanes without a ring input
    int off_win, off_yring, off_zring;
    int cy, cz, W;
};

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
// producer's step; every round polls the next four entries and forwards the valid prefix.
__device__ __forceinline__ void tile_fetch(const TileParams& P, const TileCtx& C, int f, int lane) {
    const int nyl = C.cy > 0 ? P.wz * TZ : 0;
    const int nzl = C.cz > 0 ? P.wy * TY : 0;
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
            dst = smem_u32(C.smem + (size_t)cons * P.per_warp + C.off_yring) + b * P.ring * 16;
            lead = DY;
        }
    } else if (line < nyl + nzl) {
        const int l = line - nyl;
        const int wy = l / TY, a = l % TY;
        const int A = C.cy * P.wy + wy, B = C.cz * P.wz;
        if (A < P.nty && B < P.ntz && (B % P.ntz1) != 0) {
            cons = wy;
            src = P.zmail + ((size_t)(B - 1) * P.nyp + (A * TY + a)) * P.nsteps_pad;
            dst = smem_u32(C.smem + (size_t)cons * P.per_warp + C.off_zring) + a * P.ring * 16;
            lead = DZ;
        }
    }
    const unsigned long long ehi = (P.epoch & 0xffffffffull) << 32;
    const unsigned ringmask = (unsigned)P.ring - 1u;
    int e = src ? 0 : P.nsteps_pad;  // first producer step this lane has not forwarded
    unsigned spins = 0;
    while (true) {
        // entry e overwrites entry e - ring, which the consumer reads in its step e - ring - lead
        const int limit = min(P.nsteps_pad, kSteps * ld_prog(&C.prog[cons]) + P.ring + lead);
        unsigned long long xb[4], tg[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            tg[u] = 0ull;
            xb[u] = 0ull;
            if (e + u < limit) ld_b128_gpu(src + e + u, xb[u], tg[u]);
        }
        int adv = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (adv == u && e + u < limit && tg[u] == (ehi | (unsigned long long)(e + u + 1))) {
                st_ring_s(dst + (((unsigned)(e + u) & ringmask) << 4), __longlong_as_double((long long)xb[u]), (unsigned long long)(e + u + 1));
                adv = u + 1;
            }
        }
        e += adv;
        if (__all_sync(FULL, e >= P.nsteps_pad)) break;
        TILE_SPIN(spins, 302)
    }
}

// ---------------------------------------------------------------- compute warp: one tile, one wavefront level per step
template <bool EXACT, bool BWD, bool FACTOR>
__device__ __forceinline__ void tile_compute(const CUtensorMap* tmap, const TileParams& P, const TileCtx& C, int warp, int lane) {
    const int wy = warp % P.wy, wz = warp / P.wy;
    const int A = C.cy * P.wy + wy, B = C.cz * P.wz + wz;
    if (A >= P.nty || B >= P.ntz) return;
    const int a = lane % TY, b = lane / TY;
    const int sigma = a + b, sl = sigma & 7, sh = sigma >> 3;
    const int nstages = P.nstages, nst = P.nst, nwin = P.nwin;
    const unsigned ringmask = (unsigned)P.ring - 1u;
    const int ring_stages = P.ring / kSteps;
    unsigned char* mine = C.smem + (size_t)warp * P.per_warp;
    const uint32_t mine_s = smem_u32(mine);
    const uint32_t coef_s = mine_s + lane * 16;
    const uint32_t win_s = mine_s + C.off_win;
    const uint32_t row_s = win_s + (BWD ? 31 - lane : lane) * 64;
    const int k0 = B * TZ;  // first plane of the tile (solve numbering)
    const bool has_yin = A > 0;
    const bool has_zin = (B % P.ntz1) != 0;
    const bool has_yout = A + 1 < P.nty;
    const bool has_zout = B + 1 < P.ntz && ((B + 1) % P.ntz1) != 0;
    const uint32_t dummy_s = smem_u32(C.dummy);
    const uint32_t yin_s = (a == 0 && has_yin) ? mine_s + C.off_yring + b * P.ring * 16 : dummy_s;
    const uint32_t zin_s = (b == 0 && has_zin) ? mine_s + C.off_zring + a * P.ring * 16 : dummy_s;
    // where this lane's hand-offs go: the ring of another warp of this CTA or a mailbox in global memory; the step
    // loop issues the same generic 16-byte store either way.  Entries are indexed by OUR step.
    const bool ylane = has_yout && a == TY - 1;
    const bool zlane = has_zout && b == TZ - 1;
    unsigned char* ydst0 = nullptr;
    unsigned char* zdst0 = nullptr;
    unsigned ymask = ringmask, zmask = ringmask;
    const int* ycons = C.always;  // stage counter of whoever reads a RING we write
    const int* zcons = C.always;
    if (has_yout) {
        if (wy + 1 < P.wy) {
            ycons = &C.prog[warp + 1];
            ydst0 = C.smem + (size_t)(warp + 1) * P.per_warp + C.off_yring + b * P.ring * 16;
        } else {
            ydst0 = (unsigned char*)(P.ymail + ((size_t)A * P.nzp + (k0 + b)) * P.nsteps_pad);
            ymask = 0x0fffffffu;
        }
    }
    if (has_zout) {
        if (wz + 1 < P.wz) {
            zcons = &C.prog[warp + P.wy];
            zdst0 = C.smem + (size_t)(warp + P.wy) * P.per_warp + C.off_zring + a * P.ring * 16;
        } else {
            zdst0 = (unsigned char*)(P.zmail + ((size_t)B * P.nyp + (A * TY + a)) * P.nsteps_pad);
            zmask = 0x0fffffffu;
        }
    }
    uint64_t* bar = C.mbar + warp * nst;
    const unsigned long long ehi = (P.epoch & 0xffffffffull) << 32;  // tag = {epoch, step + 1}
    const int order_xyz = P.order_xyz, unit_diag = P.unit_diag;
    const double* coef_task = P.coef + (size_t)(B * P.nty + A) * P.nsteps_pad * 128;
    // TMA box of window block m: columns [8m, 8m + 8) of the tile's lines.  Backward solves: mirrored, with the padding
    // of ragged extents first in solve order, so that no box starts at a negative coordinate
    const int c1 = BWD ? P.nyp - TY * (A + 1) : A * TY;
    const int c2 = BWD ? P.nzp - TZ * (B + 1) : B * TZ;
    unsigned spins = 0;

    auto issue_stage = [&](int qq) {  // lane 0: coefficient stage qq and window block qq -> barrier of slot qq % nst
        const int sc = qq % nst;
        const bool blk = !FACTOR && qq < P.nblk;
        mbar_arrive_expect_tx(&bar[sc], (unsigned)(kCoefBytes + (blk ? kBlockBytes : 0)));
        bulk_g2s(mine + (size_t)sc * kCoefBytes, coef_task + (size_t)qq * (kCoefBytes / 8), (unsigned)kCoefBytes, &bar[sc]);
        if (blk) tma_load_3d(win_s + (qq % nwin) * kBlockBytes, tmap, BWD ? P.nxp - kSteps * (qq + 1) : kSteps * qq, c1, c2, smem_u32(&bar[sc]));
    };
    auto store_block = [&](int m) {  // lane 0
        tma_store_3d(tmap, BWD ? P.nxp - kSteps * (m + 1) : kSteps * m, c1, c2, win_s + (m % nwin) * kBlockBytes);
    };

    // the vector -- and, right after a (re)factorisation, the coefficient stream -- belong to the previous kernels of
    // the stream until those grids have completed: nothing of this solve reads global memory before this point
    if (P.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (lane == 0)
        for (int qq = 0; qq < nst && qq < nstages; ++qq) issue_stage(qq);

    double xprev = 0.0;
    int slot = 0;
    unsigned phase = 0;
    int sA = (4 * nwin - sh) % nwin;   // window slot of block q - sh
    int sB = (sA + nwin - 1) % nwin;   // ... of block q - sh - 1
    int yseen = 0, zseen = 0, ypre = 0, zpre = 0;
    long long* trace = P.trace ? P.trace + (size_t)(B * P.nty + A) * (nstages + 1) : nullptr;
    for (int q = 0; q < nstages; ++q) {
        if (trace && lane == 0) trace[q] = global_ns();
        // flow control: the ring slots this stage overwrites have been read (progress fetched during the last stage)
        yseen = max(yseen, ypre);
        zseen = max(zseen, zpre);
        while (yseen < q - ring_stages + 1) {
            yseen = ld_prog(ycons);
            TILE_SPIN(spins, 306)
        }
        while (zseen < q - ring_stages + 1) {
            zseen = ld_prog(zcons);
            TILE_SPIN(spins, 307)
        }
        // coefficient stage q and window block q have landed
        while (!mbar_try_wait(&bar[slot], phase)) TILE_SPIN(spins, 304)
        const uint32_t st = coef_s + slot * kCoefBytes;
        const uint32_t wA = row_s + sA * kBlockBytes + (BWD ? (7 + sl) * 8 : -sl * 8);
        const uint32_t wB = row_s + sB * kBlockBytes + (BWD ? (sl - 1) * 8 : (8 - sl) * 8);
        const int s0 = q * kSteps;
        const uint32_t yb0 = yin_s + (((unsigned)s0 & ringmask) << 4), yb1 = yin_s + (((unsigned)(s0 + kSteps) & ringmask) << 4);
        const uint32_t zb0 = zin_s + (((unsigned)s0 & ringmask) << 4), zb1 = zin_s + (((unsigned)(s0 + kSteps) & ringmask) << 4);
        unsigned char* ydst = ydst0 + (size_t)((unsigned)s0 & ymask) * 16;
        unsigned char* zdst = zdst0 + (size_t)((unsigned)s0 & zmask) * 16;
        asm volatile("" : "+l"(ydst), "+l"(zdst));  // keep them in registers: no per-step re-derivation
        const bool has_next = q + 1 < nstages;      // entries of producer steps beyond the last stage do not exist
        const int ibase = s0 - sigma - P.xoff;
        // operands and inputs of the stage's first step (later steps are requested one step ahead)
        double cz_, cy_, cx_, dv, bb;
        unsigned long long vy, ty, vz, tz;
        lds_2f64(st, cz_, cy_);
        lds_2f64(st + 512, cx_, dv);
        bb = lds_f64(0 >= sl ? wA : wB);
        ld_s128(yb0 + DY * 16, vy, ty);
        ld_s128(zb0 + DZ * 16, vz, tz);
#pragma unroll
        for (int u = 0; u < kSteps; ++u) {
            // inputs of this step (polled during the previous one): producer steps s + DY / s + DZ
            const int needy = (u + DY >= kSteps && !has_next) ? 0 : s0 + u + DY + 1;
            const int needz = (u + DZ >= kSteps && !has_next) ? 0 : s0 + u + DZ + 1;
            while (!__all_sync(FULL, (int)(unsigned)ty >= needy && (int)(unsigned)tz >= needz)) {
                TILE_SPIN(spins, 305)
                ld_s128((u + DY < kSteps ? yb0 : yb1) + ((u + DY) & 7) * 16, vy, ty);
                ld_s128((u + DZ < kSteps ? zb0 : zb1) + ((u + DZ) & 7) * 16, vz, tz);
            }
            const double czu = cz_, cyu = cy_, cxu = cx_, dvu = dv, bbu = bb;
            const double ringy = __longlong_as_double((long long)vy), ringz = __longlong_as_double((long long)vz);
            const uint32_t wu = (u >= sl ? wA : wB) + (BWD ? -u * 8 : u * 8);
            if (u + 1 < kSteps) {  // next step's operands and inputs: independent of everything below
                lds_2f64(st + (u + 1) * 1024, cz_, cy_);
                lds_2f64(st + (u + 1) * 1024 + 512, cx_, dv);
                bb = lds_f64((u + 1 >= sl ? wA : wB) + (BWD ? -(u + 1) * 8 : (u + 1) * 8));
                ld_s128((u + 1 + DY < kSteps ? yb0 : yb1) + ((u + 1 + DY) & 7) * 16, vy, ty);
                ld_s128((u + 1 + DZ < kSteps ? zb0 : zb1) + ((u + 1 + DZ) & 7) * 16, vz, tz);
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
            const bool active = (unsigned)(ibase + u) < (unsigned)P.nx;
            if (active) sts_f64(wu, xnew);
            if (FACTOR) xnew = 1.0 / xnew;
            xprev = xnew;
            // ---- hand-offs: {x, {epoch, step + 1}}
            const unsigned long long tag64 = ehi | (unsigned long long)(unsigned)(s0 + u + 1);
            if (ylane) st_gen128(ydst + u * 16, xnew, tag64);
            if (zlane) st_gen128(zdst + u * 16, xnew, tag64);
            if (u == 5) {  // consumers' progress for the next stage's flow control
                ypre = ld_prog(ycons);
                zpre = ld_prog(zcons);
            }
        }
        // block q - 2 of the window is complete (the last lane finishes column 8m + 7 in step 8m + 7 + 10): make the
        // lanes' stores visible to the TMA engine and send it out; then refill the coefficient slot just consumed
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            const int mb = q - 2;
            if (mb >= 0 && mb < P.nblk) store_block(mb);
            bulk_commit();
            if (q + nst < nstages) {
                // window slot of block q + nst: its previous occupant, block q - 3, went out in the previous stage
                bulk_wait_read<1>();
                issue_stage(q + nst);
            }
            *(volatile int*)&C.prog[warp] = q + 1;
        }
        if (++slot == nst) {
            slot = 0;
            phase ^= 1u;
        }
        sB = sA;
        if (++sA == nwin) sA = 0;
    }
    if (lane == 0) {
        for (int mb = max(nstages - 2, 0); mb < P.nblk; ++mb) store_block(mb);
        bulk_commit();
        bulk_wait_all();
        if (trace) trace[nstages] = global_ns();
    }
}

template <bool EXACT, bool BWD, bool FACTOR>
__global__ void __launch_bounds__((kMaxW + kMaxFetch) * 32, 1)
    k_trsv_tile(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TileCtx C;
    C.smem = smem;
    C.W = P.wy * P.wz;
    C.off_win = P.nst * kCoefBytes;
    C.off_yring = C.off_win + P.nwin * kBlockBytes;
    C.off_zring = C.off_yring + TZ * P.ring * 16;
    C.dummy = smem + (size_t)C.W * P.per_warp;
    C.mbar = reinterpret_cast<uint64_t*>(C.dummy + P.ring * 16);
    C.prog = reinterpret_cast<int*>(C.mbar + C.W * P.nst);
    C.always = C.prog + kMaxW;
    C.cy = blockIdx.x / P.gz;
    C.cz = blockIdx.x % P.gz;
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
    if (threadIdx.x <= kMaxW) C.prog[threadIdx.x] = threadIdx.x == kMaxW ? INT_MAX : 0;
    fence_mbar_init();
    fence_proxy_async();  // the zeroed windows are written again by the TMA engine
    __syncthreads();
    // programmatic dependent launch: the next kernel of the stream (the second solve of an apply) may run its prologue
    // on idle SMs while this one runs; it touches global memory only behind griddepcontrol.wait
    if (P.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp < W) tile_compute<EXACT, BWD, FACTOR>(&tmap, P, C, warp, lane);
    else if (warp - W < P.nfetch) tile_fetch(P, C, warp - W, lane);
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
    if (first_use_on_this_device(attr_set, ki))
        ILUB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return k;
}

size_t tile_smem(int W, int nst, int nwin, int ring) {
    const size_t per_warp = (size_t)nst * kCoefBytes + (size_t)nwin * kBlockBytes + (size_t)(TY + TZ) * ring * 16;
    return (size_t)W * per_warp + (size_t)ring * 16 + (size_t)W * nst * 8 + (kMaxW + 2) * 4 + 256;
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
    for (const Cand& c : cands) {
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
    P.per_warp = Y.tnst * kCoefBytes + Y.tnwin * kBlockBytes + (TY + TZ) * Y.tring * 16;
    const int lines = (Y.tgy > 1 ? Y.twz * TZ : 0) + (Y.tgz > 1 ? Y.twy * TY : 0);
    const int groups = (lines + 31) / 32;
    P.nfetch = groups == 0 ? 0 : std::max(groups, (int)std::min<int64_t>(kMaxFetch, std::max<int64_t>(1, opt.tile_fetch)));
    P.order_xyz = Y.order_xyz;
    P.unit_diag = T.unit_diag ? 1 : 0;
    P.ymail = Y.ymail.p;
    P.zmail = Y.zmail.p;
    P.status = h->status.p;
    P.spin_iters = (unsigned)std::min<long long>(0x7fffffffll, std::max<long long>(1 << 16, (long long)opt.spin_limit_ms * h->clock_khz / 64));
    const bool bwd = fcoef ? false : T.backward;
    const tile_kern_t kern = tile_kernel(exact, bwd, fcoef != nullptr);
    const size_t trace_len = (size_t)Y.ntasks * (P.nstages + 1);
    if (opt.trace && T.trace.n < trace_len) T.trace.alloc(trace_len);
    P.trace = opt.trace ? T.trace.p : nullptr;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(Y.tgy * Y.tgz, 1, 1);
    cfg.blockDim = dim3((Y.twy * Y.twz + P.nfetch) * 32, 1, 1);
    cfg.dynamicSmemBytes = Y.smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    P.pdl = opt.sys_pdl != 0 ? 1 : 0;
    cfg.numAttrs = P.pdl ? 1 : 0;
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
