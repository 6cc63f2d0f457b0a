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
#include <algorithm>
#include <climits>

#include "ilub_common.cuh"
#include "ilub_ptx.cuh"

namespace ilub {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr long long kAbsentBits = 0x7ff8dead00000001ll;  // quiet NaN with a payload arithmetic never produces
constexpr int kStageSteps = 8;                            // steps per coefficient stage (8 KB)
constexpr int kStageBytes = kStageSteps * 4 * 32 * 8;
constexpr int kWin = 64;                                  // columns of the rhs / solution window per line
constexpr int kMaxWarps = 6;                              // compute warps per CTA
constexpr int kMaxYChan = 8;

inline int blocks_for(int64_t work, int threads) { return (int)((work + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------ analysis
// distinct |row - dependency| offsets of the rows [0, n1): list[0..3] (0 = empty), list[4] = overflow flag
__global__ void k_sys_offsets(int n1, SrcDev S, int* list) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n1) return;
    for (int q = src_beg(S, r); q < src_end(S, r); ++q) {
        int off = S.dep_idx[q] - r;
        if (off < 0) off = -off;
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
};

// every row of ONE system: where do its coefficients go in the (task, step, lane) layout?
__global__ void k_sys_map(SysGeom G, SrcDev S, int* __restrict__ srcpos, int* fail) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= G.n1) return;
    const int rm = G.backward ? G.n1 - 1 - r : r;
    const int i = rm % G.nx;
    const int j = (rm / G.nx) % G.ny;
    const int k = rm / (G.nx * G.ny);
    const int task = k * G.nyb + (j >> 5);
    const int lane = j & 31;
    const int step = i + lane;
    const size_t base = ((size_t)task * G.nsteps_pad + step) * 128 + lane;
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
        srcpos[base + role * 32] = S.val_pos ? S.val_pos[q] : q;
    }
    if (!G.unit_diag) srcpos[base + 3 * 32] = S.diag_pos[r];
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
    const int pos = srcpos[(size_t)(k * nyb + jb) * per_task + rem];
    const int role = (rem >> 5) & 3;
    double v;
    if (pos >= 0) {
        v = lu[(int64_t)pos + (int64_t)sys * nnz1];
        if (role == 3 && recip) v = 1.0 / v;
    } else {
        v = role == 3 ? 1.0 : __longlong_as_double(kAbsentBits);
    }
    coef[e] = v;
}

// ------------------------------------------------------------------------------------------------ kernel
struct SysParams {
    const double* coef;
    double* x;
    int nx, ny, nyb, nplanes, nz1, nsteps_pad;
    int pyb, pz, gyb;
    int backward;
    long long n;
    ulonglong2* zmail;
    ulonglong2* ymail;
    unsigned long long epoch;
    int zcoupled, order_xyz, unit_diag;
    int nst, rz, ry;
    int per_warp;  // shared-memory bytes per compute warp
    int* status;
    long long spin_limit;
};

struct Dog {  // clock64 watchdog: a broken schedule becomes ILUB_ERR_WATCHDOG instead of a hang
    long long t0 = 0;
    unsigned spins = 0;
    int* status;
    long long limit;
    __device__ __forceinline__ Dog(int* st, long long lim) : status(st), limit(lim) {}
    __device__ __forceinline__ bool expired(int code) {
        if ((++spins & 1023u) != 0) return false;
        if (t0 == 0) t0 = clock64();
        if (*(volatile int*)status != 0) return true;
        if (clock64() - t0 > limit) {
            atomicCAS(status, 0, code);
            return true;
        }
        return false;
    }
};

__device__ __forceinline__ int ld_prog(const int* p) { return *(const volatile int*)p; }

template <bool EXACT>
__global__ void __launch_bounds__(256, 1) k_trsv_systolic(SysParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = P.pyb * P.pz;
    const int zbytes = P.zcoupled ? P.rz * 512 : 0;
    const int ring_off = P.nst * kStageBytes + 32 * kWin * 8;  // zin ring, then yin ring
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + (size_t)W * P.per_warp);
    int* prog = reinterpret_cast<int*>(mbar + W * P.nst);

    // rings start with tag 0 (never a valid tag), barriers armed, progress 0
    for (int w = 0; w < W; ++w) {
        ulonglong2* ring = reinterpret_cast<ulonglong2*>(smem + (size_t)w * P.per_warp + ring_off);
        const int cnt = (zbytes + P.ry * 16) / 16;
        for (int t = threadIdx.x; t < cnt; t += blockDim.x) ring[t] = make_ulonglong2(0ull, 0ull);
    }
    if (threadIdx.x < W * P.nst) mbar_init(&mbar[threadIdx.x], 1);
    if (threadIdx.x < W) prog[threadIdx.x] = 0;
    fence_mbar_init();
    __syncthreads();

    const int cy = blockIdx.x % P.gyb, cz = blockIdx.x / P.gyb;
    Dog dog(P.status, P.spin_limit);

    if (warp == W) {
        // ============================================================ z fetcher: plane cz*pz - 1 -> warp 0
        // (z-coupled tilings have pyb == 1: one remote z channel per CTA)
        const int p = cz * P.pz, jb = cy * P.pyb;
        if (!P.zcoupled || p >= P.nplanes || (p % P.nz1) == 0 || jb >= P.nyb) return;
        const size_t chan = (size_t)(cz - 1) * P.nyb + jb;
        const ulonglong2* src = P.zmail + chan * P.nsteps_pad * 32 + lane;
        const uint32_t zin = smem_u32(smem + ring_off) + lane * 16;
        const int rzm = P.rz - 1;
        for (int s0 = 0; s0 < P.nsteps_pad; s0 += 4) {
            while (s0 + 3 >= ld_prog(&prog[0]) + P.rz)
                if (dog.expired(201)) return;
            bool done[4] = {false, false, false, false};
            while (true) {
                unsigned long long xb[4], tg[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (!done[u]) ld_b128_gpu(src + (size_t)(s0 + u) * 32, xb[u], tg[u]);
                bool pending = false;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!done[u]) {
                        if (tg[u] == P.epoch) {
                            st_ring_s(zin + (((s0 + u) & rzm) << 9), __longlong_as_double((long long)xb[u]),
                                      (unsigned long long)(s0 + u) + 1ull);
                            done[u] = true;
                        } else {
                            pending = true;
                        }
                    }
                }
                if (!__any_sync(FULL, pending)) break;
                if (dog.expired(202)) return;
            }
        }
        return;
    }
    if (warp == W + 1) {
        // ============================================================ y fetcher: line block cy*pyb - 1 -> lane 0 of
        // the first line block of every plane of this CTA (up to kMaxYChan channels polled together)
        const int jb = cy * P.pyb;
        if (cy == 0 || jb >= P.nyb) return;
        int i0[kMaxYChan];
        bool done[kMaxYChan];
        int nch = 0;
#pragma unroll
        for (int c = 0; c < kMaxYChan; ++c) {
            i0[c] = 0;
            done[c] = false;
            if (c < P.pz && cz * P.pz + c < P.nplanes) nch = c + 1;
        }
        const int rym = P.ry - 1;
        while (true) {
            unsigned long long xb[kMaxYChan], tg[kMaxYChan];
            bool can[kMaxYChan];
            bool active = false;
#pragma unroll
            for (int c = 0; c < kMaxYChan; ++c) {
                can[c] = false;
                if (c < nch && i0[c] < P.nx) {
                    active = true;
                    const int i = i0[c] + lane;
                    const int cons = c * P.pyb;  // warp (wy = 0, wz = c)
                    can[c] = i < P.nx && !done[c] && i < ld_prog(&prog[cons]) + P.ry;
                    if (can[c]) {
                        const size_t chan = (size_t)(cz * P.pz + c) * (P.gyb - 1) + (cy - 1);
                        ld_b128_gpu(P.ymail + chan * P.nx + i, xb[c], tg[c]);
                    }
                }
            }
            if (!active) break;
#pragma unroll
            for (int c = 0; c < kMaxYChan; ++c) {
                if (c < nch && i0[c] < P.nx) {
                    const int i = i0[c] + lane;
                    if (can[c] && tg[c] == P.epoch) {
                        const uint32_t yin =
                            smem_u32(smem + (size_t)(c * P.pyb) * P.per_warp + ring_off + zbytes) + ((i & rym) << 4);
                        st_ring_s(yin, __longlong_as_double((long long)xb[c]), (unsigned long long)i + 1ull);
                        done[c] = true;
                    }
                    if (__all_sync(FULL, done[c] || i >= P.nx)) {
                        i0[c] += 32;
                        done[c] = false;
                    }
                }
            }
            if (dog.expired(203)) return;
        }
        return;
    }

    // ================================================================ compute warps
    const int wy = warp % P.pyb, wz = warp / P.pyb;
    const int jb = cy * P.pyb + wy, p = cz * P.pz + wz;
    if (jb >= P.nyb || p >= P.nplanes) return;
    const int task = p * P.nyb + jb;
    const bool has_zin = P.zcoupled && (p % P.nz1) != 0;
    const bool has_zout = P.zcoupled && ((p + 1) % P.nz1) != 0;
    const bool zout_local = has_zout && wz + 1 < P.pz;
    const bool has_yin = jb > 0;
    const bool has_yout = jb + 1 < P.nyb;
    const bool yout_local = has_yout && wy + 1 < P.pyb;
    const int j0 = jb * 32;
    const int nlines = min(32, P.ny - j0);
    const int nblocks = P.nsteps_pad / kStageSteps;
    const int nrect = (P.nx + 7) >> 3;
    const long long tile_base = ((long long)p * P.ny + j0) * P.nx;

    unsigned char* mine = smem + (size_t)warp * P.per_warp;
    const double* stage = reinterpret_cast<const double*>(mine);
    double* xb = reinterpret_cast<double*>(mine + P.nst * kStageBytes);  // [32][kWin]
    const uint32_t zin_s = smem_u32(mine + ring_off) + lane * 16;
    const uint32_t yin_s = smem_u32(mine + ring_off + zbytes);
    const uint32_t zout_s = smem_u32(smem + (size_t)(warp + P.pyb) * P.per_warp + ring_off) + lane * 16;
    const uint32_t yout_s = smem_u32(smem + (size_t)(warp + 1) * P.per_warp + ring_off + zbytes);
    uint64_t* bar = mbar + warp * P.nst;
    const int rzm = P.rz - 1, rym = P.ry - 1;
    ulonglong2* zmail_out = nullptr;
    ulonglong2* ymail_out = nullptr;
    if (has_zout && !zout_local)
        zmail_out = P.zmail + ((size_t)((p + 1) / P.pz - 1) * P.nyb + jb) * P.nsteps_pad * 32 + lane;
    if (has_yout && !yout_local) ymail_out = P.ymail + ((size_t)p * (P.gyb - 1) + ((jb + 1) / P.pyb - 1)) * P.nx;
    const double* coef_task = P.coef + (size_t)task * P.nsteps_pad * 128;

    // rhs / solution rectangles: 32 lines x 8 columns, element e = q*32 + lane -> line e>>3, column e&7
    const int rl = lane >> 3, rc = lane & 7;
    auto row_of = [&](int line, int col) -> long long {
        const long long rm = tile_base + (long long)line * P.nx + col;
        return P.backward ? P.n - 1 - rm : rm;
    };
    double pre[8];
    auto rect_load = [&](int m) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int line = q * 4 + rl, col = m * 8 + rc;
            pre[q] = (line < nlines && col < P.nx) ? __ldcs(P.x + row_of(line, col)) : 0.0;
        }
    };
    auto rect_deposit = [&](int m) {
#pragma unroll
        for (int q = 0; q < 8; ++q) xb[(q * 4 + rl) * kWin + ((m * 8 + rc) & (kWin - 1))] = pre[q];
    };
    auto rect_drain = [&](int m) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int line = q * 4 + rl, col = m * 8 + rc;
            const double v = xb[line * kWin + (col & (kWin - 1))];
            if (line < nlines && col < P.nx) P.x[row_of(line, col)] = v;
        }
    };
    auto issue_stage = [&](int b) {  // lane 0 only
        uint64_t* bb = &bar[b % P.nst];
        mbar_arrive_expect_tx(bb, (unsigned)kStageBytes);
        bulk_g2s(mine + (size_t)(b % P.nst) * kStageBytes, coef_task + (size_t)b * (kStageBytes / 8),
                 (unsigned)kStageBytes, bb);
    };

    if (lane == 0)
        for (int b = 0; b < P.nst && b < nblocks; ++b) issue_stage(b);
    rect_load(0);
    rect_deposit(0);
    if (nrect > 1) rect_load(1);
    __syncwarp();

    double xprev = 0.0;
    int zlimit = 0, ylimit = 0;  // flow control: steps / elements the local consumer has room for
    for (int B = 0; B < nblocks; ++B) {
        if (B + 1 < nrect) rect_deposit(B + 1);
        if (B + 2 < nrect) rect_load(B + 2);
        if (B >= 5 && B - 5 < nrect) rect_drain(B - 5);
        __syncwarp();
        {
            const unsigned par = (unsigned)(B / P.nst) & 1u;
            while (!mbar_try_wait(&bar[B % P.nst], par))
                if (__any_sync(FULL, dog.expired(204))) return;
        }
        const double* st = stage + (size_t)(B % P.nst) * (kStageBytes / 8) + lane;
#pragma unroll
        for (int u = 0; u < kStageSteps; ++u) {
            const int s = B * kStageSteps + u;
            const int i = s - lane;
            const bool valid = lane < nlines && i >= 0 && i < P.nx;
            double cz_ = st[u * 128], cy_ = st[u * 128 + 32], cx_ = st[u * 128 + 64];
            const double dv = st[u * 128 + 96];
            const double bb = valid ? xb[lane * kWin + (i & (kWin - 1))] : 0.0;
            const bool az = __double_as_longlong(cz_) == kAbsentBits;
            const bool ay = __double_as_longlong(cy_) == kAbsentBits;
            const bool ax = __double_as_longlong(cx_) == kAbsentBits;
            double xy = __shfl_up_sync(FULL, xprev, 1);
            double xz = 0.0;
            // ---- inputs from other warps: z ring (all lanes), y ring (lane 0)
            const bool need_y = has_yin && lane == 0 && s < P.nx;
            {
                bool zok = !has_zin, yok = !need_y;
                const uint32_t za = zin_s + ((s & rzm) << 9);
                const uint32_t ya = yin_s + ((s & rym) << 4);
                while (true) {
                    unsigned long long vb, tg;
                    if (!zok) {
                        ld_b128_cta(za, vb, tg);
                        if ((unsigned)tg == (unsigned)(s + 1)) {
                            xz = __longlong_as_double((long long)vb);
                            zok = true;
                        }
                    }
                    if (!yok) {
                        ld_b128_cta(ya, vb, tg);
                        if ((unsigned)tg == (unsigned)(s + 1)) {
                            xy = __longlong_as_double((long long)vb);
                            yok = true;
                        }
                    }
                    if (__all_sync(FULL, zok && yok)) break;
                    if (__any_sync(FULL, dog.expired(205))) return;
                }
            }
            if (lane == 0 && !need_y) xy = 0.0;
            if (az) { cz_ = 0.0; xz = 0.0; }
            if (ay) { cy_ = 0.0; xy = 0.0; }
            const double xx = ax ? 0.0 : xprev;
            if (ax) cx_ = 0.0;
            double xnew;
            if (EXACT) {
                double acc = bb;
                if (P.order_xyz) {
                    if (!ax) acc = __dsub_rn(acc, __dmul_rn(cx_, xx));
                    if (!ay) acc = __dsub_rn(acc, __dmul_rn(cy_, xy));
                    if (!az) acc = __dsub_rn(acc, __dmul_rn(cz_, xz));
                } else {
                    if (!az) acc = __dsub_rn(acc, __dmul_rn(cz_, xz));
                    if (!ay) acc = __dsub_rn(acc, __dmul_rn(cy_, xy));
                    if (!ax) acc = __dsub_rn(acc, __dmul_rn(cx_, xx));
                }
                xnew = P.unit_diag ? acc : __ddiv_rn(acc, dv);
            } else {
                double acc = fma(-cz_, xz, bb);
                acc = fma(-cx_, xx, acc);
                acc = fma(-cy_, xy, acc);
                xnew = acc * dv;
            }
            xprev = xnew;
            if (valid) xb[lane * kWin + (i & (kWin - 1))] = xnew;
            // ---- outputs
            if (has_zout) {
                if (zout_local) {
                    while (s >= zlimit) {
                        zlimit = ld_prog(&prog[warp + P.pyb]) + P.rz;
                        if (s < zlimit) break;
                        if (__any_sync(FULL, dog.expired(206))) return;
                    }
                    st_ring_s(zout_s + ((s & rzm) << 9), xnew, (unsigned long long)s + 1ull);
                } else {
                    st_mail(zmail_out + (size_t)s * 32, xnew, P.epoch);
                }
            }
            if (has_yout && lane == 31 && valid) {
                if (yout_local) {
                    while (i >= ylimit) {
                        ylimit = ld_prog(&prog[warp + 1]) + P.ry;
                        if (i < ylimit) break;
                        if (dog.expired(207)) return;
                    }
                    st_ring_s(yout_s + ((i & rym) << 4), xnew, (unsigned long long)i + 1ull);
                } else {
                    st_mail(ymail_out + i, xnew, P.epoch);
                }
            }
            if (lane == 0) *(volatile int*)&prog[warp] = s + 1;
        }
        __syncwarp();
        if (lane == 0 && B + P.nst < nblocks) issue_stage(B + P.nst);
    }
    __syncwarp();
    for (int m = max(0, nblocks - 5); m < nrect; ++m) rect_drain(m);
}

}  // namespace

// ================================================================================================== host
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
    k_sys_offsets<<<blocks_for(n1, 256), 256, 0, s>>>(n1, S, dl.p);
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
    G.nyb = (G.ny + 31) / 32;
    G.nsteps_pad = (G.nx + 31 + kStageSteps - 1) / kStageSteps * kStageSteps;
    G.backward = src.backward ? 1 : 0;
    G.unit_diag = src.unit_diag ? 1 : 0;

    // ---- tiling of tasks onto CTAs, shared-memory budget
    const int zc = G.sz > 0 ? 1 : 0;
    const int nplanes = G.nz1 * nsys;
    const int rz = 8, ry = 64;
    const size_t budget = (size_t)227 * 1024 - 1024;
    auto per_warp = [&](int nst) { return (size_t)nst * kStageBytes + 32 * kWin * 8 + (zc ? rz * 512 : 0) + ry * 16; };
    auto max_warps = [&](int nst) { return (int)std::min<size_t>(kMaxWarps, budget / (per_warp(nst) + 64)); };
    int pyb = 1, pz = 1, nst = 3;
    if (zc) {
        pz = 1;
        while ((int64_t)G.nyb * ((nplanes + pz - 1) / pz) > h->sm_count) ++pz;
        if (opt.sys_pz > 0) pz = (int)opt.sys_pz;
        pz = std::min(pz, nplanes);
        if (pz > max_warps(2) || pz > kMaxYChan) return false;  // too large for one co-resident wave: generic path
        nst = 4;
        while (nst > 2 && pz > max_warps(nst)) --nst;
    } else {
        const int mw = max_warps(3);
        const int gyb = (G.nyb + mw - 1) / mw;
        pyb = (G.nyb + gyb - 1) / gyb;
        nst = 4;
        while (nst > 3 && pyb > max_warps(nst)) --nst;
    }
    if (opt.sys_stages >= 2 && opt.sys_stages <= 4 && pyb * pz <= max_warps((int)opt.sys_stages)) nst = (int)opt.sys_stages;

    // ---- (task, step, lane) map of one system; fails on anything that is not the grid stencil
    const size_t per_task = (size_t)G.nsteps_pad * 128;
    const int ntasks1 = G.nyb * G.nz1;
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
    Y.nx = G.nx;
    Y.ny = G.ny;
    Y.nz1 = G.nz1;
    Y.nplanes = nplanes;
    Y.nyb = G.nyb;
    Y.nsteps_pad = G.nsteps_pad;
    Y.pyb = pyb;
    Y.pz = pz;
    Y.gyb = (G.nyb + pyb - 1) / pyb;
    Y.gz = (nplanes + pz - 1) / pz;
    Y.zcoupled = zc;
    Y.order_xyz = slot == 1 ? 1 : 0;  // bwd-gather accumulates i+1, i+nx, i+nx*ny in stored order
    Y.ntasks1 = ntasks1;
    Y.ntasks = ntasks1 * nsys;
    Y.nstages = nst;
    Y.rz = rz;
    Y.ry = ry;
    Y.smem = (size_t)pyb * pz * per_warp(nst) + (size_t)pyb * pz * nst * 8 + kMaxWarps * 4 + 64;
    Y.coef.alloc((size_t)Y.ntasks * per_task);
    Y.nzchan = zc ? G.nyb * std::max(0, Y.gz - 1) : 0;
    Y.nychan = nplanes * std::max(0, Y.gyb - 1);
    Y.zmail.alloc((size_t)Y.nzchan * G.nsteps_pad * 32);
    Y.ymail.alloc((size_t)Y.nychan * G.nx);
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
    k_sys_pack<<<blocks_for(total, 256), 256, 0, s>>>(total, per_task, Y.nyb, Y.nz1, h->nsys, src.backward ? 1 : 0,
                                                      h->nnz / h->nsys, exact ? 0 : 1, Y.srcpos.p, h->lu.p, Y.coef.p);
    h->launches++;
    Y.packed_exact = exact;
}

void run_systolic(ilub_handle* h, int slot, double* x, int64_t ldx, int nrhs, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    SysSolve& Y = T.sys;
    const Options& opt = options();
    const bool exact = opt.exact != 0;
    if (Y.packed_exact != (exact ? 1 : 0)) pack_systolic(h, slot, s);
    SysParams P;
    P.coef = Y.coef.p;
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
    P.zcoupled = Y.zcoupled;
    P.order_xyz = Y.order_xyz;
    P.unit_diag = T.unit_diag ? 1 : 0;
    P.nst = Y.nstages;
    P.rz = Y.rz;
    P.ry = Y.ry;
    P.per_warp = (int)((size_t)Y.nstages * kStageBytes + 32 * kWin * 8 + (Y.zcoupled ? Y.rz * 512 : 0) + Y.ry * 16);
    P.status = h->status.p;
    P.spin_limit = (long long)opt.spin_limit_ms * h->clock_khz;
    static thread_local bool attr_set[2] = {false, false};
    if (!attr_set[exact ? 1 : 0]) {
        if (exact)
            ILUB_CUDA(cudaFuncSetAttribute(k_trsv_systolic<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        else
            ILUB_CUDA(cudaFuncSetAttribute(k_trsv_systolic<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[exact ? 1 : 0] = true;
    }
    const int grid = Y.gyb * Y.gz;
    const int block = (Y.pyb * Y.pz + 2) * 32;
    for (int c = 0; c < nrhs; ++c) {
        P.epoch = ++Y.epoch;
        P.x = x + (int64_t)c * ldx;
        if (exact)
            k_trsv_systolic<true><<<grid, block, Y.smem, s>>>(P);
        else
            k_trsv_systolic<false><<<grid, block, Y.smem, s>>>(P);
        h->launches++;
    }
    ILUB_CUDA(cudaGetLastError());
}

}  // namespace ilub
