// Line-systolic sparse triangular solves for structured grids -- the fast path behind apply() /
// apply_trans() (replaces matrix_sparse::triangular_solve, sparse_implementation.h:4040-4087) whenever the
// analysis finds that every dependency of a solve sits at offset 1, nx or nx*ny of a lexicographically
// numbered nx x ny x nplanes grid (5-/7-point stencils, their ILU(0)/IChol(0) factors, block-diagonal batches).
//
// Schedule.  A grid line (fixed j, plane p; nx rows chained by the offset-1 dependency) belongs to ONE lane,
// which walks it front to back and keeps the previous value of the line in a register.  A warp owns 32
// adjacent lines of a plane (a "task") and runs them skewed by one step per lane: at step s lane l resolves
// element i = s - l, so the offset-nx dependency (i, j-1) is what lane l-1 produced one step earlier -- one
// shuffle.  The offset-nx*ny dependency (i, j, p-1) was produced by the same lane of the warp that owns the
// previous plane at the same step: that warp writes {x, step tag} into the consumer's shared-memory ring and
// the consumer polls it with one 16-byte load.  Nothing on the dependency path touches global memory: matrix
// coefficients arrive in shared memory through 1-D bulk (TMA) copies staged three 8-step blocks ahead, the
// right-hand side is moved in and the solution out in coalesced 32 x 8 rectangles through a shared-memory
// window.  One wavefront level costs one warp step (shuffle + 2-3 dependent FMAs) instead of a flag round trip.
//
// Tasks of a CTA: pyb line blocks x pz planes, one compute warp each.  Hand-offs that cross CTAs (the first
// plane of a CTA's plane group, lane 0 of its first line block) go through global mailboxes {x, epoch}
// written with one 16-byte store; two fetcher warps per CTA poll them several steps ahead and drop the
// values into the same shared-memory rings, so compute warps never wait on an L2 round trip per step.
//
// Arithmetic: `exact` = separately rounded multiply / subtract in the reference's accumulation order and a
// true division (bit-for-bit with a non-FMA build of the reference); otherwise FMA and a reciprocal diagonal.
// Absent stencil entries are marked with a reserved NaN payload and skipped, so NaN / inf propagate exactly
// as in the reference.
#include <cooperative_groups.h>

#include <algorithm>
#include <climits>
#include <cstdlib>

#include "ilub_common.cuh"
#include "ilub_ptx.cuh"

namespace ilub {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr long long kAbsentBits = kSysAbsentBits;        // quiet NaN with a payload arithmetic never produces
constexpr int kStageSteps = kSysStageSteps;               // steps per coefficient stage (8 KB)
constexpr int kStageBytes = kStageSteps * 4 * 32 * 8;
constexpr int kMaxWarps = 6;                              // compute warps per CTA (each has a helper warp)
constexpr int kMaxYChan = 8;
// Mailbox fetchers poll global memory in blocking rounds (one L2 round trip each); a hand-off waits half a round on
// average before the next poll is even issued.  Several fetcher warps running the same loop (forwarding is
// idempotent) sample the mailbox proportionally more often.
constexpr int kZFetchWarps = 3, kYFetchWarps = 2;
inline int fetch_warps(int compute_warps) { return std::min(kZFetchWarps + kYFetchWarps, 14 - 2 * compute_warps); }

inline int blocks_for(int64_t work, int threads) { return (int)((work + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------ analysis
// distinct |row - dependency| offsets of the rows [0, n1): list[0..3] (0 = empty), list[4] = overflow flag
// (a SAMPLE of the rows -- `nruns` runs of 256 consecutive rows spread over the system, so that every grid direction
// varies inside the sample: k_sys_map below verifies every entry of every row against the offsets found here, so an offset
// that only unsampled rows have makes the map fail -- the same outcome as finding it here)
__global__ void k_sys_offsets(int n1, int nruns, SrcDev S, int* list) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int run = t >> 8;
    if (run >= nruns) return;
    const int64_t r64 = (int64_t)run * (n1 / nruns) + (t & 255);
    if (r64 >= n1) return;
    const int r = (int)r64;
    // what is on the list already (after the first few rows: everything) is recognised without touching it again
    int known[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) known[k] = *(volatile int*)(list + k);
    for (int q = src_beg(S, r); q < src_end(S, r); ++q) {
        int off = S.dep_idx[q] - r;
        if (off < 0) off = -off;
        if (off == known[0] || off == known[1] || off == known[2] || off == known[3]) continue;
        bool placed = false;
        for (int k = 0; k < 4 && !placed; ++k) {
            int cur = *(volatile int*)(list + k);
            if (cur == 0) cur = atomicCAS(list + k, 0, off);
            if (cur == 0 || cur == off) placed = true;
        }
        if (!placed) list[4] = 1;
    }
}

struct SysGeom {
    int nx, ny, nz1, nyb, nsteps_pad;
    int sy, sz;  // offsets of the y and z dependencies (sz = 0: 2-D)
    int n1;
    int backward, unit_diag;
    int ty, tz;  // lines per task along y and z (line-systolic: 32 x 1; tile-systolic: 8 x 4); nyb = tasks along y
    // tile-systolic backward solves: the mirrored grid is shifted so that the padding of ragged extents comes FIRST
    // in solve order (TMA boxes then never start at a negative coordinate of the original vector)
    int xoff, yoff, zoff;
};
// task / lane / step of grid point (i, j, k) of one system: lane = a + ty * b runs its line skewed by a + b steps
__device__ __forceinline__ void sys_place(const SysGeom& G, int i, int j, int k, int& task, int& lane, int& step) {
    const int jt = j / G.ty, a = j - jt * G.ty, kt = k / G.tz, b = k - kt * G.tz;
    task = kt * G.nyb + jt;
    lane = a + G.ty * b;
    step = i + a + b;
}

// every row of ONE system: where do its coefficients go in the (task, step, lane) layout?
__global__ void k_sys_map(SysGeom G, SrcDev S, int* __restrict__ srcpos, int* fail) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= G.n1) return;
    const int rm = G.backward ? G.n1 - 1 - r : r;
    const int i = rm % G.nx;
    const int j = (rm / G.nx) % G.ny;
    const int k = rm / (G.nx * G.ny);
    int task, lane, step;
    sys_place(G, i + G.xoff, j + G.yoff, k + G.zoff, task, lane, step);
    // per step: [lane]{z, y} (64 values) then [lane]{x, diagonal}: role r sits at (r >> 1) * 64 + lane * 2 + (r & 1)
    const size_t base = ((size_t)task * G.nsteps_pad + step) * 128 + lane * 2;
    bool bad = false;
    for (int q = src_beg(S, r); q < src_end(S, r); ++q) {
        const int dep = S.dep_idx[q];
        const int off = G.backward ? dep - r : r - dep;
        int role = -1;
        if (off == 1 && i > 0) role = 2;
        else if (off == G.sy && j > 0) role = 1;
        else if (G.sz > 0 && off == G.sz && k > 0) role = 0;
        if (role < 0) {
            bad = true;
            break;
        }
        srcpos[base + (role >> 1) * 64 + (role & 1)] = S.val_pos ? S.val_pos[q] : q;
    }
    if (!G.unit_diag) srcpos[base + 64 + 1] = S.diag_pos[r];
    if (bad) *fail = 1;
}

// coefficient values in schedule order; systems of a batch share the srcpos map of system 0
__global__ void k_sys_pack(int64_t total, int per_task, int nyb, int nz1, int nsys, int backward, int64_t nnz1,
                           int recip, const int* __restrict__ srcpos, const double* __restrict__ lu,
                           double* __restrict__ coef) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int task = (int)(e / per_task);
    const int rem = (int)(e - (int64_t)task * per_task);
    const int p = task / nyb, jb = task - p * nyb;
    const int sysm = p / nz1, k = p - sysm * nz1;
    const int sys = backward ? nsys - 1 - sysm : sysm;
    const size_t tbase = (size_t)(k * nyb + jb) * per_task;
    const int pos = srcpos[tbase + rem];
    const int role = ((rem >> 6) & 1) * 2 + (rem & 1);
    const int64_t soff = (int64_t)sys * nnz1;
    double v;
    if (recip) {
        // fast arithmetic: x = b/d - sum (c/d) x_dep  -> {c/d, ..., 1/d}; absent entries are plain zeros
        const int dpos = srcpos[tbase + (rem & ~127) + 64 + (rem & 62) + 1];
        const double dinv = dpos >= 0 ? 1.0 / lu[(int64_t)dpos + soff] : 1.0;
        v = role == 3 ? dinv : (pos >= 0 ? lu[(int64_t)pos + soff] * dinv : 0.0);
    } else {
        v = pos >= 0 ? lu[(int64_t)pos + soff] : (role == 3 ? 1.0 : __longlong_as_double(kAbsentBits));
    }
    coef[e] = v;
}

// ---- stencil factorisation (factor_stencil below): placement of the rows of ONE system in the (task, step, lane)
// layout of the forward solve, with the positions the factorisation needs: roles z / y / x = the strictly-lower
// entry and its PARTNER (the entry (k, r) of row k's strict upper part, -1 if absent), role 3 = the diagonal
__global__ void k_sys_map_factor(SysGeom G, const int* __restrict__ ptr, const int* __restrict__ idx,
                                 const int* __restrict__ dpos, int* __restrict__ fsrc, int* __restrict__ fpart, int* fail) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= G.n1) return;
    const int i = r % G.nx;
    const int j = (r / G.nx) % G.ny;
    const int k = r / (G.nx * G.ny);
    int task, lane, step;
    sys_place(G, i, j, k, task, lane, step);
    const size_t base = ((size_t)task * G.nsteps_pad + step) * 128 + lane * 2;
    bool bad = false;
    for (int q = ptr[r]; q < dpos[r]; ++q) {
        const int dep = idx[q];
        const int off = r - dep;
        int role = -1;
        if (off == 1 && i > 0) role = 2;
        else if (off == G.sy && j > 0) role = 1;
        else if (G.sz > 0 && off == G.sz && k > 0) role = 0;
        if (role < 0) {
            bad = true;
            break;
        }
        // partner: column r in the strict upper part of row dep
        int b = dpos[dep] + 1, e = ptr[dep + 1];
        while (b < e) {
            const int m = (b + e) >> 1;
            if (idx[m] < r) b = m + 1;
            else e = m;
        }
        const size_t slot = base + (role >> 1) * 64 + (role & 1);
        fsrc[slot] = q;
        fpart[slot] = (b < ptr[dep + 1] && idx[b] == r) ? b : -1;
    }
    fsrc[base + 64 + 1] = dpos[r];
    if (bad) *fail = 1;
}
// ILU(0) only: every strict-upper entry sits at offset 1, sy or sz too (then no update can hit anything but a diagonal)
__global__ void k_sys_check_upper(int n1, const int* __restrict__ ptr, const int* __restrict__ idx,
                                  const int* __restrict__ dpos, int sy, int sz, int* fail) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n1) return;
    for (int q = dpos[r] + 1; q < ptr[r + 1]; ++q) {
        const int off = idx[q] - r;
        if (off != 1 && off != sy && !(sz > 0 && off == sz)) *fail = 1;
    }
}
// coefficient stream of the factorisation: products a_ik * a_ki (IChol(0): a_ik^2) and the diagonal a_ii
__global__ void k_sys_pack_factor(int64_t total, int per_task, int nyb, int nz1, int64_t nnz1, int ichol,
                                  const int* __restrict__ fsrc, const int* __restrict__ fpart,
                                  const double* __restrict__ a, double* __restrict__ coef) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int task = (int)(e / per_task);
    const int rem = (int)(e - (int64_t)task * per_task);
    const int p = task / nyb, jb = task - p * nyb;
    const int sys = p / nz1, k = p - sys * nz1;
    const size_t tbase = (size_t)(k * nyb + jb) * per_task;
    const int pos = fsrc[tbase + rem];
    const int role = ((rem >> 6) & 1) * 2 + (rem & 1);
    const int64_t soff = (int64_t)sys * nnz1;
    double v;
    if (role == 3) {
        v = pos >= 0 ? a[(int64_t)pos + soff] : 1.0;  // (padding: pivot 1, nothing depends on it)
    } else if (pos < 0) {
        v = 0.0;
    } else if (ichol) {
        const double t = a[(int64_t)pos + soff];
        v = t * t;
    } else {
        const int pp = fpart[tbase + rem];
        v = pp >= 0 ? a[(int64_t)pos + soff] * a[(int64_t)pp + soff] : 0.0;
    }
    coef[e] = v;
}
// from the pivots to the factor on A's pattern: ILU(0): L_ik = a_ik / d_k, U = a with the pivot on the diagonal;
// IChol(0): L_ik = a_ik / sqrt(d_k), L_ii = sqrt(d_i)
// per row: the factor the strictly-lower entries of the COLUMN of that row are scaled with (1 / d, IChol: 1 / sqrt(d)) --
// one square root and one division per row instead of one of each per entry below
__global__ void k_stencil_scale(int n, const double* __restrict__ d, int ichol, double* __restrict__ rs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rs[i] = ichol ? 1.0 / sqrt(d[i]) : 1.0 / d[i];
}
// row of every entry of ONE system (8 lanes per row)
__global__ void k_stencil_rowof(int n1, const int* __restrict__ ptr, int* __restrict__ rowof) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(t >> 3), gl = (int)(t & 7);
    if (i >= n1) return;
    for (int p = ptr[i] + gl; p < ptr[i + 1]; p += 8) rowof[p] = i;
}
// the factor on A's pattern from the pivots: one thread per ENTRY (fully coalesced value streams); everything structural
// comes from the arrays of one system (rowof, idx, dpos of system 0: the systems of a batch repeat them)
__global__ void k_stencil_finish(int64_t nnz, int nnz1, int n1, const int* __restrict__ rowof, const int* __restrict__ idx,
                                 const int* __restrict__ dpos, const double* __restrict__ a, const double* __restrict__ d,
                                 const double* __restrict__ rs, double* __restrict__ lu, int ichol) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int sys = (int)(p / nnz1);
    const int p1 = (int)(p - (int64_t)sys * nnz1);
    const int r1 = __ldg(rowof + p1), dp1 = __ldg(dpos + r1);
    if (ichol && p1 > dp1) return;  // (IChol(0): the strict upper part of the array is not used)
    const int64_t base = (int64_t)sys * n1;
    double v = a[p];
    if (p1 < dp1) v *= rs[base + __ldg(idx + p1)];
    else if (p1 == dp1) v = ichol ? sqrt(d[base + r1]) : d[base + r1];
    lu[p] = v;
}

// ------------------------------------------------------------------------------------------------ kernel
struct SysParams {
    const double* coef;
    double* x;          // solution out (and right-hand side in, unless xin differs)
    const double* xin;  // right-hand side in: x itself, or another buffer -- e.g. the caller's pinned host vector (zero copy)
    int nx, ny, nyb, nplanes, nz1, nsteps_pad;
    int pyb, pz, gyb, gz, gzp, csize;  // gzp: gz padded to the cluster size; CTA = blockIdx.x = cy * gzp + cz
    int backward;
    long long n;
    ulonglong2* zmail;
    ulonglong2* ymail;
    unsigned long long epoch;
    int order_xyz, unit_diag;
    int nst, rz;
    int per_warp;  // shared-memory bytes per compute warp
    int pdl;  // launched with programmatic stream serialisation
    int* status;
    unsigned spin_iters;  // failed polls before a wait gives up (watchdog)
    long long* trace;     // debug (option trace): %globaltimer at every block start of every task, and at its end
};

__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Shared memory of one compute warp: [nst coefficient stages][rhs/solution window 32 x 65][z ring rz x 32][y ring 64]
// Window: element (line l, column i) lives at xw[l*65 + ((i + l) & 63)], i.e. at the STEP it is resolved in, so a
// lane's accesses inside an 8-step block are base + immediate; the odd row stride keeps them conflict-free.
// Coefficient stage: per step 1 KB = [lane]{z, y} (512 B) then [lane]{x, diagonal} (512 B): two 16-byte loads.
constexpr int kWinStride = 65;
constexpr int kWinBytes = 32 * kWinStride * 8;
constexpr int kRy = 64;
constexpr int kCtlInts = 4 * kMaxWarps + 4;  // progh[W] | hdep[W] | zcredit[W] | scratch[W] | always = INT_MAX

__device__ __forceinline__ int ld_prog(const int* p) { return *(const volatile int*)p; }
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds_2f64(uint32_t a, double& v0, double& v1) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a));
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ void give_up(int* status, int code) { atomicCAS(status, 0, code); }
__device__ __forceinline__ void ld_s128(uint32_t a, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.cta.shared.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}" : "=l"(lo), "=l"(hi) : "r"(a));
}
// one 16-byte store {x, tag} through a GENERIC address: the same instruction feeds a ring in this CTA's shared
// memory, a ring in another CTA of the cluster (distributed shared memory) or a mailbox in global memory
__device__ __forceinline__ void st_gen128(void* p, double x, unsigned long long tag) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.gpu.b128 [%0], q;\n\t}" ::"l"(p),
                 "l"(__double_as_longlong(x)), "l"(tag)
                 : "memory");
}
__device__ __forceinline__ void st_gen_s32(int* p, int v) { asm volatile("st.relaxed.gpu.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// A wait that has failed `limit` polls (or sees another warp's failure) gives up: ILUB_ERR_WATCHDOG, not a hang.
#define SYS_SPIN(spins, code)                                          \
    if ((++(spins) & 0xfffu) == 0) {                                   \
        if ((spins) >= P.spin_iters || *(volatile int*)P.status != 0) { \
            give_up(P.status, code);                                   \
            return;                                                    \
        }                                                              \
    }

struct SysCtx {  // where things live in this CTA's shared memory
    unsigned char* smem;
    uint64_t* mbar;
    int* progh;    // [W] completed half-blocks (4 steps) of every compute warp
    int* hdep;     // [W] rectangles the helper of every compute warp has deposited
    int* zcredit;  // [W] progress of a z consumer that lives in the NEXT CTA of the cluster (it pushes it here)
    int* scratch;  // [W] sink for progress pushes nobody needs
    int* always;   // INT_MAX: "consumer" that never blocks
    int W, zbytes, win_off, ring_off, cy, cz, crank;
};

// ---------------------------------------------------------------- z fetcher: last plane of CTA cz-1 (outside this
// cluster) -> warp 0.  z-coupled tilings have pyb == 1: one remote z channel per CTA; four steps in flight per lane
template <bool ZC>
__device__ __forceinline__ void role_zfetch(const SysParams& P, const SysCtx& C, int lane) {
    const int p = C.cz * P.pz, jb = C.cy * P.pyb;
    if (!ZC || p >= P.nplanes || (p % P.nz1) == 0 || jb >= P.nyb || C.crank != 0) return;
    const unsigned long long ehi = (P.epoch & 0xffffffffull) << 32;
    const int zlag = P.rz / 4 - 1;
    const size_t chan = (size_t)(C.cz - 1) * P.nyb + jb;
    const ulonglong2* src = P.zmail + chan * P.nsteps_pad * 32 + lane;
    const uint32_t zin = smem_u32(C.smem + C.ring_off) + lane * 16;
    const int rzm = P.rz - 1;
    // Sliding window: every round polls the next four entries this lane has not forwarded yet and forwards the
    // valid prefix, so whatever has arrived since the last round moves on at once (a fixed group of four that must
    // complete before the next one starts costs two L2 round trips per four steps -- slower than the producer).
    unsigned spins = 0;
    int e = 0;  // first step this lane has not forwarded
    while (true) {
        // entry s may overwrite ring slot s & (rz - 1) once the consumer has passed half-block (s >> 2) - rz / 4
        const int limit = min(P.nsteps_pad, 4 * (ld_prog(&C.progh[0]) + zlag + 1));
        unsigned long long xb[4], tg[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            tg[u] = 0ull;
            if (e + u < limit) ld_b128_gpu(src + (size_t)(e + u) * 32, xb[u], tg[u]);
        }
        int adv = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (adv == u && e + u < limit && tg[u] == (ehi | (unsigned long long)(e + u + 1))) {
                st_ring_s(zin + (((e + u) & rzm) << 9), __longlong_as_double((long long)xb[u]), tg[u]);
                adv = u + 1;
            }
        }
        e += adv;
        if (__all_sync(FULL, e >= P.nsteps_pad)) break;
        SYS_SPIN(spins, 202)
    }
}

// ---------------------------------------------------------------- y fetcher: line block cy*pyb - 1 -> the first line
// block of every plane of this CTA (up to kMaxYChan channels polled together).  Entries are indexed by the
// PRODUCER's step; lanes take 32 consecutive steps of a channel.
__device__ __forceinline__ void role_yfetch(const SysParams& P, const SysCtx& C, int lane) {
    const int jb = C.cy * P.pyb;
    if (C.cy == 0 || jb >= P.nyb) return;
    const unsigned long long ehi = (P.epoch & 0xffffffffull) << 32;
    // lane l relays the entries of steps l, l + 32, l + 64, ... of every channel, each lane at its own pace
    int sp[kMaxYChan];
    int nch = 0;
#pragma unroll
    for (int c = 0; c < kMaxYChan; ++c) {
        sp[c] = lane;
        if (c < P.pz && C.cz * P.pz + c < P.nplanes) nch = c + 1;
    }
    unsigned spins = 0;
    while (true) {
        unsigned long long xb[kMaxYChan], tg[kMaxYChan];
        bool can[kMaxYChan];
        bool active = false;
#pragma unroll
        for (int c = 0; c < kMaxYChan; ++c) {
            can[c] = false;
            if (c < nch && sp[c] < P.nsteps_pad) {
                active = true;
                const int cons = c * P.pyb;  // warp (wy = 0, wz = c)
                // the slot's previous occupant (step sp - 64) is read by the consumer at its step sp - 96
                can[c] = sp[c] < 4 * ld_prog(&C.progh[cons]) + 96;
                if (can[c]) {
                    const size_t chan = (size_t)(C.cz * P.pz + c) * (P.gyb - 1) + (C.cy - 1);
                    ld_b128_gpu(P.ymail + chan * P.nsteps_pad + sp[c], xb[c], tg[c]);
                }
            }
        }
        if (!__any_sync(FULL, active)) break;
#pragma unroll
        for (int c = 0; c < kMaxYChan; ++c) {
            if (can[c] && tg[c] == (ehi | (unsigned long long)(sp[c] + 1))) {
                const uint32_t yin =
                    smem_u32(C.smem + (size_t)(c * P.pyb) * P.per_warp + C.ring_off + C.zbytes) + ((sp[c] & (kRy - 1)) << 4);
                st_ring_s(yin, __longlong_as_double((long long)xb[c]), tg[c]);
                sp[c] += 32;
            }
        }
        SYS_SPIN(spins, 203)
    }
}

// ---------------------------------------------------------------- helper of compute warp w (same scheduler): moves the
// right-hand side in and the solution out in coalesced 32 x 8 rectangles through the window, and keeps the TMA
// coefficient stages filled.  Iteration k runs beside the compute warp's block k.
__device__ __forceinline__ void role_helper(const SysParams& P, const SysCtx& C, int w, int lane) {
    const int wy = w % P.pyb, wz = w / P.pyb;
    const int jb = C.cy * P.pyb + wy, p = C.cz * P.pz + wz;
    if (jb >= P.nyb || p >= P.nplanes) return;
    const int nx = P.nx, nst = P.nst;
    const int j0 = jb * 32;
    const int nlines = min(32, P.ny - j0);
    const int nblocks = P.nsteps_pad / kStageSteps;
    const int nrect = (nx + 7) >> 3;
    const long long tile_base = ((long long)p * P.ny + j0) * nx;
    unsigned char* mine = C.smem + (size_t)w * P.per_warp;
    const uint32_t xw_s = smem_u32(mine + C.win_off);
    uint64_t* bar = C.mbar + w * nst;
    const double* coef_task = P.coef + (size_t)(p * P.nyb + jb) * P.nsteps_pad * 128;
    // element q of a lane = (line 4q + rl, column 8m + rc)
    const int rl = lane >> 3, rc = lane & 7;
    const long long dir = P.backward ? -1 : 1;
    const long long g0 = P.backward ? P.n - 1 - (tile_base + (long long)rl * nx + rc) : tile_base + (long long)rl * nx + rc;
    double* const gp = P.x + g0;
    const double* const gin = P.xin + g0;
    const long long qstride = dir * 4 * nx;
    const uint32_t xw_rl = xw_s + rl * (kWinStride * 8);
    const bool full_lines = nlines == 32;
    double pre[8];
    auto rect_load = [&](int m) {
        const double* b = gin + dir * 8 * m;
        if (full_lines && 8 * m + 8 <= nx) {
#pragma unroll
            for (int q = 0; q < 8; ++q) pre[q] = __ldcs(b + q * qstride);
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) pre[q] = (q * 4 + rl < nlines && m * 8 + rc < nx) ? __ldcs(b + q * qstride) : 0.0;
        }
    };
    auto rect_deposit = [&](int m) {
        const int t0 = 8 * m + rc + rl;
#pragma unroll
        for (int q = 0; q < 8; ++q) sts_f64(xw_rl + q * (4 * kWinStride * 8) + (((t0 + 4 * q) & 63) << 3), pre[q]);
    };
    auto rect_drain = [&](int m) {
        const int t0 = 8 * m + rc + rl;
        double* b = gp + dir * 8 * m;
        const bool guard = !(full_lines && 8 * m + 8 <= nx);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const double v = lds_f64(xw_rl + q * (4 * kWinStride * 8) + (((t0 + 4 * q) & 63) << 3));
            if (!guard || (q * 4 + rl < nlines && m * 8 + rc < nx)) __stcs(b + q * qstride, v);
        }
    };
    auto issue_stage = [&](int b) {  // lane 0 only
        const int slot = b % nst;
        mbar_arrive_expect_tx(&bar[slot], (unsigned)kStageBytes);
        bulk_g2s(mine + (size_t)slot * kStageBytes, coef_task + (size_t)b * (kStageBytes / 8), (unsigned)kStageBytes, &bar[slot]);
    };

    // the vector -- and, right after a (re)factorisation, the coefficient stream -- belong to the previous kernels of
    // the stream until those grids have completed: nothing of this solve reads global memory before this point
    if (P.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (lane == 0)
        for (int b = 0; b < nst && b < nblocks; ++b) issue_stage(b);
    // rectangles 0 and 1 before the compute warp may start; 2 stays in registers until iteration 0
    rect_load(0);
    rect_deposit(0);
    if (nrect > 1) {
        rect_load(1);
        rect_deposit(1);
    }
    if (nrect > 2) rect_load(2);
    __syncwarp();
    fence_acq_rel_cta();
    if (lane == 0) *(volatile int*)&C.hdep[w] = 2;

    unsigned spins = 0;
    const int last = max(nblocks, nrect + 4);
    for (int k = 0; k <= last; ++k) {
        // the compute warp has finished block k - 1
        const int need = 2 * min(k, nblocks);
        while (ld_prog(&C.progh[w]) < need) SYS_SPIN(spins, 209)
        fence_acq_rel_cta();
        if (lane == 0 && k >= 1 && k - 1 + nst < nblocks) issue_stage(k - 1 + nst);  // its stage slot is free again
        if (k + 2 < nrect) rect_deposit(k + 2);  // window columns of rectangle k - 6, drained in iteration k - 1
        if (k + 3 < nrect) rect_load(k + 3);
        if (k >= 5 && k - 5 < nrect) rect_drain(k - 5);  // complete since block k - 1
        __syncwarp();
        fence_acq_rel_cta();
        if (lane == 0) *(volatile int*)&C.hdep[w] = k + 3;
    }
}

// ---------------------------------------------------------------- compute warp: one task, one wavefront level per step
template <bool EXACT, bool ZC, bool FACTOR>
__device__ __forceinline__ void role_compute(const SysParams& P, const SysCtx& C, int warp, int lane) {
    namespace cg = cooperative_groups;
    const int W = C.W;
    const int wy = warp % P.pyb, wz = warp / P.pyb;
    const int jb = C.cy * P.pyb + wy, p = C.cz * P.pz + wz;
    if (jb >= P.nyb || p >= P.nplanes) return;
    const int nx = P.nx;
    const bool has_zin = ZC && (p % P.nz1) != 0;
    const bool has_zout = ZC && ((p + 1) % P.nz1) != 0;
    const bool zout_local = has_zout && wz + 1 < P.pz;
    const bool zout_dsmem = has_zout && !zout_local && C.crank + 1 < P.csize;  // next CTA of this cluster
    const bool has_yin = jb > 0;
    const bool has_yout = jb + 1 < P.nyb;
    const bool yout_local = has_yout && wy + 1 < P.pyb;
    const int nblocks = P.nsteps_pad / kStageSteps;
    unsigned char* mine = C.smem + (size_t)warp * P.per_warp;
    const uint32_t stage_s = smem_u32(mine) + lane * 16;
    const uint32_t xw_lane = smem_u32(mine + C.win_off) + lane * (kWinStride * 8);
    const uint32_t zin_s = smem_u32(mine + C.ring_off) + lane * 16;
    const uint32_t yin_s = smem_u32(mine + C.ring_off + C.zbytes);
    // where this warp's hand-offs go: a ring of another warp of this CTA, the ring of warp wy of the next CTA of
    // the cluster, or a mailbox in global memory; the step loop issues the same generic 16-byte store either way
    unsigned char* zdst0 = nullptr;
    unsigned char* ydst0 = nullptr;
    const int* zcons = C.always;  // progress word of whoever reads a RING we write (never blocks for mailboxes)
    const int* ycons = C.always;
    if (zout_local) {
        zdst0 = C.smem + (size_t)(warp + P.pyb) * P.per_warp + C.ring_off + lane * 16;
        zcons = &C.progh[warp + P.pyb];
    } else if (zout_dsmem) {
        unsigned char* mine_ring = C.smem + (size_t)wy * P.per_warp + C.ring_off + lane * 16;
        zdst0 = (unsigned char*)cg::this_cluster().map_shared_rank((void*)mine_ring, C.crank + 1);
        zcons = &C.zcredit[wy];
    } else if (has_zout) {
        zdst0 = (unsigned char*)(P.zmail + ((size_t)((p + 1) / P.pz - 1) * P.nyb + jb) * P.nsteps_pad * 32 + lane);
    }
    const bool zring = zout_local || zout_dsmem;
    if (yout_local) {
        ydst0 = C.smem + (size_t)(warp + 1) * P.per_warp + C.ring_off + C.zbytes;
        ycons = &C.progh[warp + 1];
    } else if (has_yout) {
        ydst0 = (unsigned char*)(P.ymail + ((size_t)p * (P.gyb - 1) + ((jb + 1) / P.pyb - 1)) * P.nsteps_pad);
    }
    const bool ypred = has_yout && lane == 31;
    // a z producer in the PREVIOUS CTA of the cluster needs our progress: pushed into its zcredit word
    int* credit = &C.scratch[warp];
    if (has_zin && wz == 0 && C.crank > 0) credit = (int*)cg::this_cluster().map_shared_rank((void*)&C.zcredit[wy], C.crank - 1);
    uint64_t* bar = C.mbar + warp * P.nst;
    const unsigned rzmask = (unsigned)P.rz - 1u;
    const int zlag = P.rz / 4 - 1;  // a producer may run this many half-blocks ahead of its z consumer
    const unsigned long long ehi = (P.epoch & 0xffffffffull) << 32;  // tag = {epoch, step + 1}
    const int order_xyz = P.order_xyz, unit_diag = P.unit_diag, nst = P.nst;
    unsigned spins = 0;

    // lane 0's offset-nx dependency of element i is produced by lane 31 of the previous line block at ITS step
    // i + 31.  It is polled one step early (slot of producer step s + 32 while resolving step s: same slot
    // alignment as everything else in the block) and carried in a register; element 0 is fetched here.
    double ycur = 0.0;
    if (has_yin) {
        while (true) {
            unsigned long long v, t;
            ld_s128(yin_s + (31 << 4), v, t);
            ycur = __longlong_as_double((long long)v);
            if ((int)(unsigned)t >= 32) break;
            SYS_SPIN(spins, 208)
        }
    }

    double xprev = 0.0;
    int slot = 0;
    unsigned phase = 0;
    int zseen = 0, yseen = 0, hseen = 0;  // last progress read from the z / y ring consumers and from the helper
    long long* trace = P.trace ? P.trace + (size_t)(p * P.nyb + jb) * 3 * (nblocks + 1) : nullptr;
    long long wait_cycles = 0;
    for (int B = 0; B < nblocks; ++B) {
        long long c0 = 0;
        if (trace) {
            if (lane == 0) {
                trace[B] = global_ns();
                trace[(nblocks + 1) + B] = spins;
                trace[2 * (nblocks + 1) + B] = wait_cycles;
            }
            c0 = clock64();
        }
        {   // rectangle B (all columns <= 8B + 7) is in the window
            const int need = B + 1;  // also: every rectangle aliased by this block's window columns is drained
            while (hseen < need) {
                hseen = ld_prog(&C.hdep[warp]);
                SYS_SPIN(spins, 210)
            }
        }
        while (!mbar_try_wait(&bar[slot], phase)) SYS_SPIN(spins, 204)
        if (trace) wait_cycles += clock64() - c0;
        const uint32_t st = stage_s + slot * kStageBytes;
        const uint32_t wb = xw_lane + ((B & 7) << 6);
        const uint32_t zb = zin_s + (((unsigned)(B * kStageSteps) & rzmask) << 9);
        const uint32_t yb = yin_s + (((B + 4) & 7) << 7);  // slots of producer steps 8B + 32 ...
        const bool ypoll = B + 4 < nblocks;                 // ... which exist until four blocks before the end
        unsigned char* zdst = zring ? zdst0 + (((unsigned)(B * kStageSteps) & rzmask) << 9) : zdst0 + (size_t)B * (kStageSteps * 512);
        unsigned char* ydst = yout_local ? ydst0 + ((B & 7) << 7) : ydst0 + (size_t)B * (kStageSteps * 16);
        asm volatile("" : "+l"(zdst), "+l"(ydst));  // keep them in registers: no per-step re-derivation
        const int s0 = B * kStageSteps;
        // operands of the block's first step (later steps are loaded one step ahead)
        double cz_, cy_, cx_, dv, bb;
        lds_2f64(st, cz_, cy_);
        lds_2f64(st + 512, cx_, dv);
        bb = lds_f64(wb);
#pragma unroll
        for (int u = 0; u < kStageSteps; ++u) {
            const int tag = s0 + u + 1;
            if ((u & 3) == 0) {
                // flow control once per half-block: the ring slots we are about to overwrite have been read
                const int h = 2 * B + (u >> 2);
                while (zseen < h - zlag) {
                    zseen = ld_prog(zcons);
                    SYS_SPIN(spins, 206)
                }
                while (yseen < h - 23) {
                    yseen = ld_prog(ycons);
                    SYS_SPIN(spins, 207)
                }
            }
            const double czu = cz_, cyu = cy_, cxu = cx_, dvu = dv, bbu = bb;
            if (u + 1 < kStageSteps) {  // next step's operands: independent of everything below
                lds_2f64(st + (u + 1) * 1024, cz_, cy_);
                lds_2f64(st + (u + 1) * 1024 + 512, cx_, dv);
                bb = lds_f64(wb + (u + 1) * 8);
            }
            // ---- inputs from other warps: z ring (own entry per lane, this step), y ring (one entry read by every
            //      lane, next step).  Rings of inputs this warp does not have are pre-filled "always ready".
            unsigned long long vz = 0ull, tz = (unsigned long long)INT_MAX, vy = 0ull, ty = (unsigned long long)INT_MAX;
            if (ZC) ld_s128(zb + u * 512, vz, tz);  // first poll goes out before anything else of the step
            if (ypoll) ld_s128(yb + u * 16, vy, ty);
            double xy = __shfl_up_sync(FULL, xprev, 1);
            if (lane == 0) xy = ycur;
            double xnew, ynext;
            if (EXACT) {
                while (!__all_sync(FULL, (int)(unsigned)tz >= tag && (int)(unsigned)ty >= tag + 32)) {
                    SYS_SPIN(spins, 205)
                    if (ZC) ld_s128(zb + u * 512, vz, tz);
                    if (ypoll) ld_s128(yb + u * 16, vy, ty);
                }
                const double xz = __longlong_as_double((long long)vz);
                const bool az = !ZC || __double_as_longlong(czu) == kAbsentBits;
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
                // coefficients are pre-divided by the diagonal (dvu = 1/d); absent entries carry 0.0 and every operand
                // that can reach them is finite.  Only ONE fused multiply-add waits for the other warp's value, and it
                // is issued speculatively: the vote on the tags resolves beside it.
                // FACTOR (stencil ILU(0) / IChol(0), see factor_stencil below): the same recurrence on the pivots --
                // d = a_ii - sum p_k / d_k with p_k = a_ik * a_ki packed in the coefficient slots and a_ii in the
                // diagonal slot; what travels between rows is r = 1 / d, what is stored is d
                double acc = FACTOR ? dvu : bbu * dvu;
                acc = fma(-cxu, xprev, acc);
                acc = fma(-cyu, xy, acc);
                while (true) {
                    xnew = ZC ? fma(-czu, __longlong_as_double((long long)vz), acc) : acc;
                    if (__all_sync(FULL, (int)(unsigned)tz >= tag && (int)(unsigned)ty >= tag + 32)) break;
                    SYS_SPIN(spins, 205)
                    if (ZC) ld_s128(zb + u * 512, vz, tz);
                    if (ypoll) ld_s128(yb + u * 16, vy, ty);
                }
            }
            ynext = __longlong_as_double((long long)vy);
            if (FACTOR) {
                sts_f64(wb + u * 8, xnew);  // the pivot
                xnew = 1.0 / xnew;          // its reciprocal goes to the rows that depend on it
            } else {
                sts_f64(wb + u * 8, xnew);
            }
            xprev = xnew;
            ycur = ynext;
            // ---- hand-offs: {x, {epoch, step + 1}}
            const unsigned long long tag64 = ehi | (unsigned long long)(unsigned)tag;
            if (has_zout) st_gen128(zdst + u * 512, xnew, tag64);
            if (ypred) st_gen128(ydst + u * 16, xnew, tag64);
            if ((u & 3) == 3 && lane == 0) {
                const int hb = 2 * B + (u >> 2) + 1;
                *(volatile int*)&C.progh[warp] = hb;
                st_gen_s32(credit, hb);
            }
        }
        if (++slot == nst) {
            slot = 0;
            phase ^= 1u;
        }
    }
    if (trace && lane == 0) {
        trace[nblocks] = global_ns();
        trace[(nblocks + 1) + nblocks] = spins;
        trace[2 * (nblocks + 1) + nblocks] = wait_cycles;
    }
}

template <bool EXACT, bool ZC, bool FACTOR = false>
__global__ void __launch_bounds__(448, 1) k_trsv_systolic(const __grid_constant__ SysParams P) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    SysCtx C;
    C.smem = smem;
    C.W = P.pyb * P.pz;
    C.zbytes = ZC ? P.rz * 512 : 0;
    C.win_off = P.nst * kStageBytes;
    C.ring_off = C.win_off + kWinBytes;  // z ring, then y ring
    C.mbar = reinterpret_cast<uint64_t*>(smem + (size_t)C.W * P.per_warp);
    int* ctl = reinterpret_cast<int*>(C.mbar + C.W * P.nst);
    C.progh = ctl;
    C.hdep = ctl + kMaxWarps;
    C.zcredit = ctl + 2 * kMaxWarps;
    C.scratch = ctl + 3 * kMaxWarps;
    C.always = ctl + 4 * kMaxWarps;
    C.cy = blockIdx.x / P.gzp;
    C.cz = blockIdx.x % P.gzp;
    C.crank = (int)(blockIdx.x % P.csize);  // gzp is a multiple of the cluster size
    const int W = C.W;

    // windows zeroed; rings zeroed (tag 0 = nothing yet) -- or, for a warp WITHOUT that input, pre-filled with
    // {0.0, INT_MAX}: always ready, value zero, so the step loop needs no "do I have this input" branches
    for (int w = 0; w < W; ++w) {
        unsigned char* base = smem + (size_t)w * P.per_warp;
        ulonglong2* win = reinterpret_cast<ulonglong2*>(base + C.win_off);
        for (int t = threadIdx.x; t < kWinBytes / 16; t += blockDim.x) win[t] = make_ulonglong2(0ull, 0ull);
        const int jbw = C.cy * P.pyb + w % P.pyb, pw = C.cz * P.pz + w / P.pyb;
        const unsigned long long ztag = (ZC && (pw % P.nz1) != 0) ? 0ull : (unsigned long long)INT_MAX;
        const unsigned long long ytag = jbw > 0 ? 0ull : (unsigned long long)INT_MAX;
        ulonglong2* zr = reinterpret_cast<ulonglong2*>(base + C.ring_off);
        for (int t = threadIdx.x; t < C.zbytes / 16; t += blockDim.x) zr[t] = make_ulonglong2(0ull, ztag);
        ulonglong2* yr = reinterpret_cast<ulonglong2*>(base + C.ring_off + C.zbytes);
        for (int t = threadIdx.x; t < kRy; t += blockDim.x) yr[t] = make_ulonglong2(0ull, ytag);
    }
    if (threadIdx.x < W * P.nst) mbar_init(&C.mbar[threadIdx.x], 1);
    if (threadIdx.x < kCtlInts) ctl[threadIdx.x] = threadIdx.x == 4 * kMaxWarps ? INT_MAX : 0;
    fence_mbar_init();
    __syncthreads();
    cg::this_cluster().sync();  // every CTA of the cluster has initialised its rings before anyone writes into them
    // programmatic dependent launch: the next kernel of the stream (the second solve of an apply) may start its own
    // prologue -- shared-memory set-up, cluster sync, first coefficient stages -- on idle SMs while this one runs;
    // it touches the vector only behind griddepcontrol.wait (role_helper), i.e. after this grid has completed
    if (P.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (C.cz < P.gz) {
        if (warp < W) role_compute<EXACT, ZC, FACTOR>(P, C, warp, lane);
        else if (warp < 2 * W) role_helper(P, C, warp - W, lane);
        else {
            // fetchers: z and y alternate (z first) so that a short budget still has one of each
            const int f = warp - 2 * W;
            if ((f & 1) == 0 && f / 2 < kZFetchWarps) role_zfetch<ZC>(P, C, lane);
            else role_yfetch(P, C, lane);
        }
    }
    // nobody leaves while a neighbour may still write into its shared memory
    cg::this_cluster().sync();
}

}  // namespace

// ================================================================================================== host
typedef void (*sys_kern_t)(const SysParams);
static sys_kern_t sys_factor_kernel(bool zc) {
    const sys_kern_t k = zc ? (sys_kern_t)k_trsv_systolic<false, true, true> : (sys_kern_t)k_trsv_systolic<false, false, true>;
    static thread_local unsigned attr_set[2] = {0, 0};
    if (first_use_on_this_device(attr_set, zc ? 1 : 0)) {
        ILUB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ILUB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    return k;
}
static sys_kern_t sys_kernel(bool exact, bool zc) {
    const sys_kern_t k = exact ? (zc ? (sys_kern_t)k_trsv_systolic<true, true> : (sys_kern_t)k_trsv_systolic<true, false>)
                               : (zc ? (sys_kern_t)k_trsv_systolic<false, true> : (sys_kern_t)k_trsv_systolic<false, false>);
    static thread_local unsigned attr_set[4] = {0, 0, 0, 0};
    const int ki = (exact ? 2 : 0) + (zc ? 1 : 0);
    if (first_use_on_this_device(attr_set, ki)) {
        ILUB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ILUB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    return k;
}
// how many clusters of `csize` CTAs of this shape can be resident at once (0: not launchable)
static int max_active_clusters(sys_kern_t k, int csize, int block, size_t smem) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(csize, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize;
    at[0].val.clusterDim.y = 1;
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

bool analyse_systolic(ilub_handle* h, int slot, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    SysSolve& Y = T.sys;
    if (Y.tried) return Y.built;
    Y.tried = true;
    const Options& opt = options();
    const SolveSource src = solve_source(h, slot);
    const SrcDev S = to_dev(src);
    const int nsys = h->nsys;
    const int n1 = h->n / nsys;
    const int64_t nnz1 = h->nnz / nsys;

    DevBuf<int> dl;
    dl.alloc(8);
    ILUB_CUDA(cudaMemsetAsync(dl.p, 0, sizeof(int) * 8, s));
    const int nruns = std::max(1, std::min(32, n1 / 256));
    k_sys_offsets<<<nruns, 256, 0, s>>>(n1, nruns, S, dl.p);
    h->launches++;
    int hl[8];
    ILUB_CUDA(cudaMemcpyAsync(hl, dl.p, sizeof(hl), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    if (hl[4]) return false;
    int off[4], cnt = 0;
    for (int k = 0; k < 4; ++k)
        if (hl[k] > 0) off[cnt++] = hl[k];
    if (cnt < 2 || cnt > 3) return false;
    std::sort(off, off + cnt);
    if (off[0] != 1) return false;
    SysGeom G;
    G.n1 = n1;
    G.nx = off[1];
    G.sy = off[1];
    if (cnt == 3) {
        if (off[2] % off[1] != 0 || n1 % off[2] != 0) return false;
        G.sz = off[2];
        G.ny = off[2] / off[1];
        G.nz1 = n1 / off[2];
    } else {
        if (n1 % off[1] != 0) return false;
        G.sz = 0;
        G.ny = n1 / off[1];
        G.nz1 = 1;
    }
    if (G.nx < 2 || G.ny < 2) return false;
    G.ty = 32;
    G.tz = 1;
    G.xoff = G.yoff = G.zoff = 0;
    G.nyb = (G.ny + 31) / 32;
    G.nsteps_pad = (G.nx + 31 + kStageSteps - 1) / kStageSteps * kStageSteps;
    G.backward = src.backward ? 1 : 0;
    G.unit_diag = src.unit_diag ? 1 : 0;
    Y.nx = G.nx;
    Y.ny = G.ny;
    Y.nz1 = G.nz1;
    Y.nplanes = G.nz1 * nsys;
    Y.zcoupled = G.sz > 0 ? 1 : 0;
    Y.tile = 0;
    Y.ty = 32;
    Y.tz = 1;
    Y.ntz1 = G.nz1;
    // plane-coupled 3-D grids: the tile-systolic kernel (ilub_tile.cu) when the grid fits it
    if (Y.zcoupled && opt.sys_tile != 0 && tile_configure(h, Y, nsys)) {
        G.ty = Y.ty;
        G.tz = Y.tz;
        G.nyb = Y.nty;
        const int nxp = (G.nx + kStageSteps - 1) / kStageSteps * kStageSteps;
        G.nsteps_pad = (nxp + 2 * kStageSteps + 2 * kStageSteps - 1) / (2 * kStageSteps) * (2 * kStageSteps);  // >= nxp + ty + tz - 2, whole 16-step iterations
        if (G.backward) {
            G.xoff = nxp - G.nx;
            G.yoff = Y.nty * Y.ty - G.ny;
            G.zoff = Y.ntz1 * Y.tz - G.nz1;
        }
    }

    // ---- tiling of tasks onto CTAs, shared-memory budget
    const int zc = G.sz > 0 ? 1 : 0;
    const int nplanes = G.nz1 * nsys;
    const size_t budget = (size_t)227 * 1024 - 1024;
    auto per_warp = [&](int nst, int rz) { return (size_t)nst * kStageBytes + kWinBytes + (zc ? rz * 512 : 0) + kRy * 16; };
    auto max_warps = [&](int nst, int rz) { return (int)std::min<size_t>(kMaxWarps, budget / (per_warp(nst, rz) + 64)); };
    int pyb = 1, pz = 1, nst = 3, rz = 16, csize = 1, gzp = 1;
    const bool tile = Y.tile != 0;
    auto fit_rings = [&](int warps, int& nst_, int& rz_) {  // three stages and the 16-step z ring if they fit
        if (warps <= max_warps(3, 16)) { nst_ = 3; rz_ = 16; }
        else if (warps <= max_warps(3, 8)) { nst_ = 3; rz_ = 8; }
        else if (warps <= max_warps(2, 16)) { nst_ = 2; rz_ = 16; }
        else { nst_ = 2; rz_ = 8; }
    };
    auto smem_of = [&](int warps, int nst_, int rz_) { return (size_t)warps * per_warp(nst_, rz_) + (size_t)warps * nst_ * 8 + kCtlInts * 4 + 64; };
    if (tile) {
        // (CTA tiling chosen by tile_configure)
    } else if (zc) {
        // planes per CTA: the fewest that keep the whole grid co-resident, then up to two more if that lets a
        // larger thread-block cluster (fewer global-memory hand-offs on the z chain) stay co-resident
        int pz0 = 1;
        while ((int64_t)G.nyb * ((nplanes + pz0 - 1) / pz0) > h->sm_count) ++pz0;
        if (opt.sys_pz > 0) pz0 = (int)opt.sys_pz;
        pz0 = std::min(pz0, nplanes);
        if (pz0 > max_warps(2, 8) || pz0 > kMaxYChan) return false;  // too large for one co-resident wave: generic path
        const sys_kern_t k = sys_kernel(false, true);
        ILUB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        ILUB_CUDA(cudaFuncSetAttribute(sys_kernel(true, true), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        const int cmax = (int)std::min<int64_t>(16, opt.sys_cluster > 0 ? opt.sys_cluster : 16);
        long best_cost = LONG_MAX;
        const long long key[4] = {G.nyb, nplanes, pz0, cmax};
        const bool cached = std::equal(key, key + 4, h->sys_tiling_key);
        if (cached) {
            pz = h->sys_tiling[0], nst = h->sys_tiling[1], rz = h->sys_tiling[2], csize = h->sys_tiling[3], gzp = h->sys_tiling[4];
            best_cost = 0;
        }
        for (int pzc = pz0; !cached && pzc <= std::min(pz0 + (opt.sys_pz > 0 ? 0 : 2), std::min(nplanes, std::min(max_warps(2, 8), kMaxYChan))); ++pzc) {
            int nstc, rzc;
            fit_rings(pzc, nstc, rzc);
            const int gz = (nplanes + pzc - 1) / pzc;
            const int block = (pzc * 2 + fetch_warps(pzc)) * 32;
            const size_t sm = smem_of(pzc, nstc, rzc);
            for (int c = cmax; c >= 1; --c) {
                const int gzpc = (gz + c - 1) / c * c;
                if ((int64_t)G.nyb * gzpc > h->sm_count) continue;
                if (c > 1 && G.nyb * gzpc / c > max_active_clusters(k, c, block, sm)) continue;
                if (c == 1) {  // no cluster: plain occupancy says whether the whole grid is resident at once
                    int occ = 0;
                    ILUB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, block, sm));
                    if ((int64_t)occ * h->sm_count < (int64_t)G.nyb * gzpc) continue;
                }
                // cost model (ns on the z chain): global hop 1500, DSMEM hop 300, local hop 120; two compute warps
                // on one scheduler slow every step of the task
                const long glob = gzpc / c - 1, dsm = gz - 1 - glob, loc = nplanes - gz;
                // (measured on 128^3: a fifth / sixth compute warp per SM costs more than the hops it saves)
                const long cost = glob * 1500 + dsm * 300 + loc * 120 + (long)(pzc - pz0) * 15000 + (nstc < 3 ? 1000 : 0);
                if (getenv("ILUB_DEBUG"))
                    fprintf(stderr, "[ilub] tiling pz %d cluster %d: gz %d (padded %d), global %ld dsmem %ld, cost %ld\n", pzc, c, gz,
                            gzpc, glob, dsm, cost);
                if (cost < best_cost) {
                    best_cost = cost;
                    pz = pzc;
                    nst = nstc;
                    rz = rzc;
                    csize = c;
                    gzp = gzpc;
                }
                break;  // largest admissible cluster for this pz
            }
        }
        if (best_cost == LONG_MAX) return false;
        std::copy(key, key + 4, h->sys_tiling_key);
        h->sys_tiling[0] = pz, h->sys_tiling[1] = nst, h->sys_tiling[2] = rz, h->sys_tiling[3] = csize, h->sys_tiling[4] = gzp;
    } else {
        const int mw = max_warps(3, 8);
        const int gyb = (G.nyb + mw - 1) / mw;
        pyb = (G.nyb + gyb - 1) / gyb;
        nst = 4;
        while (nst > 3 && pyb > max_warps(nst, 8)) --nst;
    }
    if (opt.sys_stages >= 2 && opt.sys_stages <= 4 && pyb * pz <= max_warps((int)opt.sys_stages, rz)) nst = (int)opt.sys_stages;

    // ---- (task, step, lane) map of one system; fails on anything that is not the grid stencil
    const size_t per_task = (size_t)G.nsteps_pad * 128;
    const int ntasks1 = G.nyb * (tile ? Y.ntz1 : G.nz1);
    ILUB_REQUIRE((int64_t)ntasks1 * nsys * (int64_t)per_task < ((int64_t)1 << 40), ILUB_ERR_SIZE, "systolic layout too large");
    Y.srcpos.alloc((size_t)ntasks1 * per_task);
    ILUB_CUDA(cudaMemsetAsync(Y.srcpos.p, 0xff, sizeof(int) * (size_t)ntasks1 * per_task, s));
    ILUB_CUDA(cudaMemsetAsync(dl.p, 0, sizeof(int) * 8, s));
    k_sys_map<<<blocks_for(n1, 256), 256, 0, s>>>(G, S, Y.srcpos.p, dl.p);
    h->launches++;
    ILUB_CUDA(cudaMemcpyAsync(hl, dl.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    if (hl[0]) {
        Y.srcpos.release();
        return false;
    }
    Y.nyb = G.nyb;
    Y.nsteps_pad = G.nsteps_pad;
    Y.pyb = pyb;
    Y.pz = pz;
    Y.gyb = (G.nyb + pyb - 1) / pyb;
    Y.gz = (nplanes + pz - 1) / pz;
    Y.order_xyz = slot == 1 ? 1 : 0;  // bwd-gather accumulates i+1, i+nx, i+nx*ny in stored order
    Y.ntasks1 = ntasks1;
    Y.ntasks = ntasks1 * nsys;
    Y.nstages = nst;
    Y.rz = rz;
    Y.ry = kRy;
    if (!tile) Y.smem = smem_of(pyb * pz, nst, rz);
    // thread-block cluster along z (chosen with the tiling above): hand-offs between consecutive CTAs of a cluster
    // go through distributed shared memory instead of a global mailbox
    Y.csize = zc ? csize : 1;
    Y.gzp = zc ? gzp : Y.gz;
    Y.coef.alloc((size_t)Y.ntasks * per_task);
    if (tile) {
        // global mailboxes for hand-offs between CTAs: one line of nsteps_pad entries per (tile-row boundary, plane) /
        // (tile-column boundary, grid line), indexed by the producer's step
        Y.nychan = std::max(0, Y.nty - 1) * Y.ntz * Y.tz;
        Y.nzchan = std::max(0, Y.ntz - 1) * Y.nty * Y.ty;
        Y.ymail.alloc((size_t)Y.nychan * G.nsteps_pad);
        Y.zmail.alloc((size_t)Y.nzchan * G.nsteps_pad);
    } else {
        Y.nzchan = zc ? G.nyb * std::max(0, Y.gz - 1) : 0;
        Y.nychan = nplanes * std::max(0, Y.gyb - 1);
        Y.zmail.alloc((size_t)Y.nzchan * G.nsteps_pad * 32);
        Y.ymail.alloc((size_t)Y.nychan * G.nsteps_pad);
    }
    ILUB_CUDA(cudaMemsetAsync(Y.zmail.p, 0, Y.zmail.bytes(), s));
    ILUB_CUDA(cudaMemsetAsync(Y.ymail.p, 0, Y.ymail.bytes(), s));
    Y.epoch = 0;
    Y.packed_exact = -1;
    Y.built = true;
    (void)nnz1;
    return true;
}

void pack_systolic(ilub_handle* h, int slot, cudaStream_t s) {
    SysSolve& Y = h->solve[slot].sys;
    if (!Y.built) return;
    const SolveSource src = solve_source(h, slot);
    const int exact = options().exact != 0 ? 1 : 0;
    const int per_task = Y.nsteps_pad * 128;
    const int64_t total = (int64_t)Y.ntasks * per_task;
    k_sys_pack<<<blocks_for(total, 256), 256, 0, s>>>(total, per_task, Y.nyb, Y.ntz1, h->nsys, src.backward ? 1 : 0,
                                                      h->nnz / h->nsys, exact ? 0 : 1, Y.srcpos.p, h->lu.p, Y.coef.p);
    h->launches++;
    Y.packed_exact = exact;
}

static void launch_systolic(ilub_handle* h, int slot, double* x, int64_t ldx, int nrhs, cudaStream_t s, const double* fcoef,
                            const double* xin = nullptr) {
    TriSolve& T = h->solve[slot];
    SysSolve& Y = T.sys;
    const Options& opt = options();
    const bool exact = fcoef ? false : opt.exact != 0;
    if (!fcoef && Y.packed_exact != (exact ? 1 : 0)) pack_systolic(h, slot, s);
    SysParams P;
    P.coef = fcoef ? fcoef : Y.coef.p;
    P.nx = Y.nx;
    P.ny = Y.ny;
    P.nyb = Y.nyb;
    P.nplanes = Y.nplanes;
    P.nz1 = Y.nz1;
    P.nsteps_pad = Y.nsteps_pad;
    P.pyb = Y.pyb;
    P.pz = Y.pz;
    P.gyb = Y.gyb;
    P.backward = T.backward ? 1 : 0;
    P.n = h->n;
    P.zmail = Y.zmail.p;
    P.ymail = Y.ymail.p;
    P.order_xyz = Y.order_xyz;
    P.unit_diag = T.unit_diag ? 1 : 0;
    P.nst = Y.nstages;
    P.rz = Y.rz;
    P.per_warp = (int)((size_t)Y.nstages * kStageBytes + kWinBytes + (Y.zcoupled ? Y.rz * 512 : 0) + kRy * 16);
    P.status = h->status.p;
    // watchdog in failed polls (a poll round is >= ~40 cycles)
    P.spin_iters = (unsigned)std::min<long long>(0x7fffffffll, std::max<long long>(1 << 16, (long long)opt.spin_limit_ms * h->clock_khz / 64));
    const sys_kern_t kern = fcoef ? sys_factor_kernel(Y.zcoupled != 0) : sys_kernel(exact, Y.zcoupled != 0);
    if (opt.trace && T.trace.n < (size_t)Y.ntasks * 3 * (Y.nsteps_pad / kStageSteps + 1))
        T.trace.alloc((size_t)Y.ntasks * 3 * (Y.nsteps_pad / kStageSteps + 1));
    P.trace = opt.trace ? T.trace.p : nullptr;
    P.gz = Y.gz;
    P.gzp = Y.gzp;
    P.csize = Y.csize;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(Y.gyb * Y.gzp, 1, 1);
    cfg.blockDim = dim3((Y.pyb * Y.pz * 2 + fetch_warps(Y.pyb * Y.pz)) * 32, 1, 1);
    cfg.dynamicSmemBytes = Y.smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = Y.csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    P.pdl = opt.sys_pdl != 0 ? 1 : 0;
    cfg.numAttrs = P.pdl ? 2 : 1;
    for (int c = 0; c < nrhs; ++c) {
        P.epoch = ++Y.epoch;
        P.x = x + (int64_t)c * ldx;
        P.xin = xin ? xin + (int64_t)c * ldx : P.x;
        ILUB_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
        h->launches++;
    }
    ILUB_CUDA(cudaGetLastError());
}

// out-of-place solve of one right-hand side on the line-systolic kernel: xin -> xout, either of which may be the caller's
// pinned host vector (the kernel then streams it over PCIe while the wavefront advances: no staging copy)
bool systolic_out_of_place_ok(const ilub_handle* h, int slot) {
    const SysSolve& Y = h->solve[slot].sys;
    return h->solve[slot].built && Y.built && !Y.tile && options().solve_algo == 0;
}
void run_systolic_io(ilub_handle* h, int slot, const double* xin, double* xout, cudaStream_t s) {
    launch_systolic(h, slot, xout, h->n, 1, s, nullptr, xin);
}

void run_systolic(ilub_handle* h, int slot, double* x, int64_t ldx, int nrhs, cudaStream_t s) {
    if (run_plane(h, slot, x, ldx, nrhs, s)) return;
    if (h->solve[slot].sys.tile) {
        if (h->solve[slot].sys.packed_exact != (options().exact != 0 ? 1 : 0)) pack_systolic(h, slot, s);
        launch_tile(h, slot, x, ldx, nrhs, s, nullptr);
    } else {
        launch_systolic(h, slot, x, ldx, nrhs, s, nullptr);
    }
}

// ================================================================================================== stencil factorisation
// ILU(0) / IChol(0) of a 5- / 7-point stencil matrix change nothing but the pivots: for the lower neighbours k of
// row i (offsets 1, nx, nx*ny) no entry (k, j), j != i, of row k's upper part meets an entry of row i once nx >= 3
// and ny >= 3, so the only update is  d_i = a_ii - sum_k a_ik a_ki / d_k  (IChol(0): d_i = l_ii^2, a_ik^2 / d_k).
// That is the wavefront recurrence of the forward solve with a reciprocal in it, and it runs on the same
// line-systolic kernel (template parameter FACTOR): the coefficient stream carries the products, the vector carries
// the pivots.  Replaces compute_ilu0 / compute_ichol0 (ILU0.hpp:8-66, IChol.hpp:15-59) for such matrices when `exact`
// is off: the result differs from the reference's operation order by the rounding of (a_ik a_ki) / d_k against
// (a_ik / d_k) a_ki -- a few 1e-16 relative (tests hold it to 1e-12); `exact` keeps the general kernels.
bool stencil_factor_possible(ilub_handle* h, cudaStream_t s) {
    SysSolve& Y = h->solve[0].sys;
    if (Y.ftried) return Y.fbuilt;
    Y.ftried = true;
    if (!Y.built || Y.nx < 3 || (Y.zcoupled && Y.ny < 3)) return false;
    const int n1 = h->n / h->nsys;
    DevBuf<int> fail;
    fail.alloc(1);
    ILUB_CUDA(cudaMemsetAsync(fail.p, 0, sizeof(int), s));
    SysGeom G;
    G.n1 = n1;
    G.nx = Y.nx;
    G.ny = Y.ny;
    G.nz1 = Y.nz1;
    G.nyb = Y.nyb;
    G.nsteps_pad = Y.nsteps_pad;
    G.sy = Y.nx;
    G.sz = Y.zcoupled ? Y.nx * Y.ny : 0;
    G.backward = 0;
    G.unit_diag = 0;
    G.ty = Y.ty;
    G.tz = Y.tz;
    G.xoff = G.yoff = G.zoff = 0;
    const size_t per_task = (size_t)Y.nsteps_pad * 128;
    Y.fsrc.alloc((size_t)Y.ntasks1 * per_task);
    Y.fpart.alloc((size_t)Y.ntasks1 * per_task);
    ILUB_CUDA(cudaMemsetAsync(Y.fsrc.p, 0xff, sizeof(int) * (size_t)Y.ntasks1 * per_task, s));
    ILUB_CUDA(cudaMemsetAsync(Y.fpart.p, 0xff, sizeof(int) * (size_t)Y.ntasks1 * per_task, s));
    // (the pattern of a batch is that of its first system repeated: rows [0, n1) of the expanded arrays)
    k_sys_map_factor<<<blocks_for(n1, 256), 256, 0, s>>>(G, h->ptr.p, h->idx.p, h->dpos.p, Y.fsrc.p, Y.fpart.p, fail.p);
    if (h->kind != KIND_ICHOL0)
        k_sys_check_upper<<<blocks_for(n1, 256), 256, 0, s>>>(n1, h->ptr.p, h->idx.p, h->dpos.p, G.sy, G.sz, fail.p);
    h->launches += 2;
    int hf = 0;
    ILUB_CUDA(cudaMemcpyAsync(&hf, fail.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    if (hf) {
        Y.fsrc.release();
        Y.fpart.release();
        return false;
    }
    Y.fcoef.alloc((size_t)Y.ntasks * per_task);
    h->fpivot.alloc(h->n);
    Y.fbuilt = true;
    return true;
}

void factor_stencil(ilub_handle* h, cudaStream_t s) {
    SysSolve& Y = h->solve[0].sys;
    const int ichol = h->kind == KIND_ICHOL0 ? 1 : 0;
    const int per_task = Y.nsteps_pad * 128;
    const int64_t total = (int64_t)Y.ntasks * per_task;
    k_sys_pack_factor<<<blocks_for(total, 256), 256, 0, s>>>(total, per_task, Y.nyb, Y.ntz1, h->nnz / h->nsys, ichol, Y.fsrc.p,
                                                             Y.fpart.p, h->a.p, Y.fcoef.p);
    if (Y.tile) launch_tile(h, 0, h->fpivot.p, h->n, 1, s, Y.fcoef.p);
    else launch_systolic(h, 0, h->fpivot.p, h->n, 1, s, Y.fcoef.p);
    h->work.ensure((size_t)h->n);  // (vector workspace: free between applies)
    k_stencil_scale<<<blocks_for(h->n, 256), 256, 0, s>>>(h->n, h->fpivot.p, ichol, h->work.p);
    const int n1 = h->n / h->nsys;
    const int nnz1 = (int)(h->nnz / h->nsys);
    if (Y.frow.n < (size_t)nnz1) {
        Y.frow.alloc((size_t)nnz1);
        k_stencil_rowof<<<blocks_for((int64_t)n1 * 8, 256), 256, 0, s>>>(n1, h->ptr.p, Y.frow.p);
        h->launches++;
    }
    k_stencil_finish<<<blocks_for(h->nnz, 256), 256, 0, s>>>(h->nnz, nnz1, n1, Y.frow.p, h->idx.p, h->dpos.p, h->a.p, h->fpivot.p,
                                                            h->work.p, h->lu.p, ichol);
    h->launches += 3;
    ILUB_CUDA(cudaGetLastError());
}

}  // namespace ilub
