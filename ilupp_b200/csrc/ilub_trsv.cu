// Sparse triangular solves behind apply() / apply_trans() -- replaces the four sequential loops of
// matrix_sparse::triangular_solve (sparse_implementation.h:4040-4087).
//
// Sync-free kernel (default).  One persistent CTA per row partition, all CTAs co-resident.  The
// CTA's rows are packed in (level, row) order into 32-row chunks; warp w takes chunks w, w+W, ...
// A chunk's index/value slabs are staged into shared memory ahead of time with 1-D bulk async
// copies (TMA engine, mbarrier completion) so that only the x-gather sits on the dependency
// critical path.  Every lane owns one row and walks its dependencies in the reference's
// accumulation order in a non-blocking "try-advance" loop:
//   * dependency in the same chunk        -> value passed lane to lane with a shuffle
//   * dependency in a recent chunk of the
//     same CTA (<= near chunks back)      -> x ring + per-chunk tag in shared memory
//                                            (st.release.cta / ld.acquire.cta)
//   * anything else (other CTA, far back) -> 16-byte mailbox {x, epoch tag} in global memory, written
//                                            and polled with single atomic .b128 accesses (no fence)
// Forward progress: chunks are ordered by (level, row) inside a CTA and every warp takes its
// chunks in ascending order, so the globally smallest unfinished row (by level, row) is always at
// the head of a resident warp and has all dependencies satisfied.  A clock64 watchdog turns any
// violation of that argument into ILUB_ERR_WATCHDOG instead of a hang.
//
// Level-scheduled kernel (option solve_algo = 1): one launch per level over the CSR-read arrays.
#include <algorithm>
#include <climits>
#include <cub/cub.cuh>

#include "ilub_common.cuh"
#include "ilub_ptx.cuh"

namespace ilub {
namespace {

constexpr unsigned FULL = 0xffffffffu;

struct TrsvParams {
    const int* part_chunk_ptr;
    const int* chunk_ptr;
    const unsigned char* ckind;
    const int* rows;
    const int* rlen;
    const int* idx;
    const double* val;
    const double* diag;
    const unsigned* pubmask;
    const int* gate;
    int slack;
    double* x;
    ulonglong2* mail;
    unsigned long long epoch;
    int ring_mask;  // ring rows - 1
    int rc;         // ring chunks
    int near;
    int nstages;
    int stage_slabs;
    int unit_diag;
    int* status;
    long long spin_limit;
    int fastpath;
    long long* trace;  // debug: 4 clock64 stamps per chunk (start, gate passed, data staged, resolved)
};

struct Watchdog {
    long long t0 = 0;
    unsigned spins = 0;
    int* status;
    long long limit;
    __device__ __forceinline__ Watchdog(int* st, long long lim) : status(st), limit(lim) {}
    // true => give up (another CTA failed or we ran out of time)
    __device__ __forceinline__ bool expired(int code) {
        if ((++spins & 127u) != 0) return false;
        if (t0 == 0) t0 = clock64();
        if (*(volatile int*)status != 0) return true;
        if (clock64() - t0 > limit) {
            atomicCAS(status, 0, code);
            return true;
        }
        return false;
    }
    // unconditional variant for callers that already rate-limit
    __device__ __forceinline__ bool expired_now(int code) {
        if (t0 == 0) t0 = clock64();
        if (*(volatile int*)status != 0) return true;
        if (clock64() - t0 > limit) {
            atomicCAS(status, 0, code);
            return true;
        }
        return false;
    }
};

// one dependency poll: mailbox (global, e >= 0) or x ring (shared, e < 0); true when the value is there
__device__ __forceinline__ bool poll_dep(int e, const ulonglong2* mail, unsigned long long epoch,
                                         const ulonglong2* ring, int ring_mask, double& xv) {
    if (e >= 0) return ld_mail(mail + e, epoch, xv);
    if ((unsigned)e & kReadyBit) {  // remote value already folded into the staged coefficient
        xv = 1.0;
        return true;
    }
    const int lp = e & 0x3fffffff;
    return ld_ring(ring + (lp & ring_mask), (unsigned long long)lp + 1ull, xv);
}

// Critical path of a single-level lane-per-row chunk: NW ring polls in flight per lane (absent / pre-folded
// dependencies point at two always-ready dummy entries), one vote, NW multiply-subtracts in stored order.
template <int NW, bool EXACT>
__device__ __forceinline__ bool fast_resolve(const uint32_t (&a)[4], const unsigned (&g)[4], const double (&cv)[4],
                                             double& acc, Watchdog& wd) {
    unsigned spins = 0;
    double y[NW];
    while (true) {
        unsigned long long xb[NW], u[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) ld_b128_cta(a[k], xb[k], u[k]);
        bool ok = true;
#pragma unroll
        for (int k = 0; k < NW; ++k) ok = ok && ((unsigned)u[k] == g[k]);
#pragma unroll
        for (int k = 0; k < NW; ++k) y[k] = __longlong_as_double((long long)xb[k]);
        if (__all_sync(0xffffffffu, ok)) break;
        if ((++spins & 0xfffu) == 0 && wd.expired_now(107)) return false;
    }
#pragma unroll
    for (int k = 0; k < NW; ++k) acc = sub_mul<EXACT>(acc, cv[k], y[k]);
    return true;
}

template <bool EXACT>
__global__ void __launch_bounds__(512, 1) k_trsv_syncfree(TrsvParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = P.ring_mask + 1;
    ulonglong2* ring = reinterpret_cast<ulonglong2*>(smem);  // {x bits, local position + 1}
    unsigned char* stage_base = smem + (size_t)(R + 8) * 16;  // 8 spare entries: [R] = {0.0, ~0}, [R+1] = {1.0, ~0}
    const int stage_bytes = P.stage_slabs * 384;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(stage_base + (size_t)W * P.nstages * stage_bytes);
    int* warp_cur = reinterpret_cast<int*>(mbar + W * P.nstages);

    const int c0 = P.part_chunk_ptr[blockIdx.x];
    const int nloc = P.part_chunk_ptr[blockIdx.x + 1] - c0;
    for (int i = threadIdx.x; i < R; i += blockDim.x) ring[i] = make_ulonglong2(0ull, 0ull);
    if (threadIdx.x == 0) {
        ring[R] = make_ulonglong2(0ull, 0xffffffffffffffffull);
        ring[R + 1] = make_ulonglong2((unsigned long long)__double_as_longlong(1.0), 0xffffffffffffffffull);
    }
    if (threadIdx.x < 32) warp_cur[threadIdx.x] = (threadIdx.x < W && threadIdx.x < nloc) ? threadIdx.x : INT_MAX;
    if (lane == 0) {
        for (int s = 0; s < P.nstages; ++s) mbar_init(&mbar[warp * P.nstages + s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (warp >= nloc) return;

    unsigned phasebits = 0;
    int done_cache = 0;
    Watchdog wd(P.status, P.spin_limit);

    // lane 0: start the bulk copies of local chunk lc into this warp's stage (it % nstages)
    auto issue = [&](int lc, int it) {
        const int c = c0 + lc;
        const int cp = P.chunk_ptr[c];
        const int w = P.chunk_ptr[c + 1] - cp;
        if ((P.ckind[c] & 1) == 0 && w > 0) {
            const int st = it % P.nstages;
            unsigned char* sb = stage_base + (size_t)(warp * P.nstages + st) * stage_bytes;
            uint64_t* bar = &mbar[warp * P.nstages + st];
            mbar_arrive_expect_tx(bar, (unsigned)w * 384u);
            bulk_g2s(sb, P.val + (size_t)cp * 32, (unsigned)w * 256u, bar);
            bulk_g2s(sb + (size_t)P.stage_slabs * 256, P.idx + (size_t)cp * 32, (unsigned)w * 128u, bar);
        }
    };

    // software pipeline over this warp's chunks: chunk metadata and the right-hand side of the NEXT chunk
    // are loaded while the current one is being resolved
    int lc = warp;
    if (lane == 0) issue(lc, 0);
    int r = P.rows[(size_t)(c0 + lc) * 32 + lane];
    int len = P.rlen[(size_t)(c0 + lc) * 32 + lane];
    double d = P.unit_diag ? 1.0 : P.diag[(size_t)(c0 + lc) * 32 + lane];
    double acc = r >= 0 ? __ldcg(P.x + r) : 0.0;

    for (int it = 0; lc < nloc; lc += W, ++it) {
        const int c = c0 + lc;
        const int nlc = lc + W;
        const bool has_next = nlc < nloc;
        fence_proxy_async();  // generic-proxy writes into a stage (remote prefetch) before the next bulk copy into it
        __syncwarp();
        if (has_next && lane == 0 && P.nstages > 1) issue(nlc, it + 1);
        int r1 = -1, len1 = 0;
        double d1 = 1.0;
        if (has_next) {
            r1 = P.rows[(size_t)(c0 + nlc) * 32 + lane];
            len1 = P.rlen[(size_t)(c0 + nlc) * 32 + lane];
            if (!P.unit_diag) d1 = P.diag[(size_t)(c0 + nlc) * 32 + lane];
        }
        if (P.trace && lane == 0) P.trace[(size_t)c * 4 + 0] = clock64();
        const int cp = P.chunk_ptr[c];
        const int w = P.chunk_ptr[c + 1] - cp;
        const unsigned kind = P.ckind[c];
        const unsigned pubm = P.pubmask[c];
        const bool pubm_remote = (kind & 4) != 0;  // chunk has mailbox dependencies
        double xfin = 0.0;

        // (1) ring safety: the ring slots of chunk lc are those of chunk lc - rc; they may be overwritten once
        //     every chunk that could still read them through shared memory (ids <= lc - rc + near) is done.
        const int need = lc - P.rc + P.near + 1;
        if (need > done_cache) {
            while (true) {
                int v = lane < W ? ld_relaxed_cta_s32(&warp_cur[lane]) : INT_MAX;
                v = __reduce_min_sync(FULL, v);
                done_cache = v;
                if (v >= need) break;
                __nanosleep(64);
                if (wd.expired(105)) return;
            }
        }
        // (2) gate row: the last row (in processing order) that feeds this chunk's lowest level.  Until it is
        //     published nothing in this chunk can finish, so the whole warp waits on that ONE ring entry
        //     (same address for all lanes: a broadcast, one shared-memory wavefront per poll) instead of
        //     32 lanes polling 32 different entries and saturating the shared-memory pipe for the warps
        //     that are on the critical path.  The gate row has a strictly lower level than every row of
        //     this chunk, so blocking on it cannot create a cycle.
        const int gate_lp = P.gate[c];
        auto wait_gate = [&]() -> bool {
            if (gate_lp < 0 || P.slack < 0) return true;
            const uint32_t ga = smem_u32(ring + (gate_lp & P.ring_mask));
            const unsigned gt = (unsigned)gate_lp + 1u;
            const int gate_chunk = gate_lp >> 5;
            // far from the frontier: watch the per-warp progress words and sleep
            while (gate_chunk - done_cache > P.slack) {
                int v = lane < W ? ld_relaxed_cta_s32(&warp_cur[lane]) : INT_MAX;
                done_cache = __reduce_min_sync(FULL, v);
                if (gate_chunk - done_cache <= P.slack) break;
                __nanosleep(100);
                if (wd.expired(106)) return false;
            }
            // near: one lane polls the gate row (one shared-memory wavefront per round)
            unsigned spins = 0;
            while (true) {
                bool open = false;
                if (lane == 0) {
                    unsigned long long xb, t;
                    ld_b128_cta(ga, xb, t);
                    open = ((unsigned)t == gt);
                }
                if (__any_sync(FULL, open)) return true;
                if ((++spins & 0xfffu) == 0 && wd.expired_now(106)) return false;
            }
        };
        if (P.trace && lane == 0) P.trace[(size_t)c * 4 + 1] = clock64();
        ulonglong2* myslot = ring + ((lc * 32 + lane) & P.ring_mask);
        const unsigned long long mytag = (unsigned long long)(lc * 32 + lane) + 1ull;

        if ((kind & 1) == 0) {
            // ------------------------------------------------------------ lane-per-row chunk
            const int st = it % P.nstages;
            const unsigned char* sb = stage_base + (size_t)(warp * P.nstages + st) * stage_bytes;
            const double* sval = reinterpret_cast<const double*>(sb);
            const int* sidx = reinterpret_cast<const int*>(sb + (size_t)P.stage_slabs * 256);
            if (w > 0) {
                uint64_t* bar = &mbar[warp * P.nstages + st];
                const unsigned par = (phasebits >> st) & 1u;
                while (!mbar_try_wait(bar, par))
                    if (wd.expired(102)) return;
                phasebits ^= 1u << st;
            }
            const double dinv = EXACT ? 0.0 : 1.0 / d;
            // Remote prefetch: resolve mailbox dependencies up front (a few bounded attempts, all lanes and
            // slabs in parallel) and fold them into the staged coefficient (v := v*x, index := READY), so
            // that the ordered hand-off loop below never waits on an L2 round trip in the middle of a
            // chain.  Anything not there yet is simply polled in place later -- no blocking, no deadlock.
            bool remote_pending = false;
            if (pubm_remote) {
                int* widx = const_cast<int*>(sidx);
                double* wval = const_cast<double*>(sval);
                for (int attempt = 0; attempt < 4; ++attempt) {
                    bool pending = false;
                    for (int jj = 0; jj < len; ++jj) {
                        const int ee = widx[jj * 32 + lane];
                        if (ee >= 0) {
                            double xr;
                            if (ld_mail(P.mail + ee, P.epoch, xr)) {
                                const double vv = wval[jj * 32 + lane];
                                wval[jj * 32 + lane] = EXACT ? __dmul_rn(vv, xr) : vv * xr;
                                widx[jj * 32 + lane] = (int)(kLocalBit | kReadyBit);
                            } else {
                                pending = true;
                            }
                        }
                    }
                    remote_pending = __any_sync(FULL, pending);
                    if (!remote_pending) break;
                }
            }
            // the first (up to) four dependencies of the lane go to registers while the warp is still waiting
            // at its gate: once the gate opens they are polled all at once (one shared-memory round trip)
            int q0 = 0, q1 = 0, q2 = 0, q3 = 0;
            double c0v = 0.0, c1v = 0.0, c2v = 0.0, c3v = 0.0;
            if (len > 0) { q0 = sidx[lane];       c0v = sval[lane]; }
            if (len > 1) { q1 = sidx[32 + lane];  c1v = sval[32 + lane]; }
            if (len > 2) { q2 = sidx[64 + lane];  c2v = sval[64 + lane]; }
            if (len > 3) { q3 = sidx[96 + lane];  c3v = sval[96 + lane]; }
            // everything the fast path needs is computed BEFORE the gate: what runs after it is the critical path
            uint32_t fa[4];
            unsigned fg[4];
            double fc[4];
            {
                const uint32_t rb = smem_u32(ring);
                const int qq[4] = {q0, q1, q2, q3};
                const double cc[4] = {c0v, c1v, c2v, c3v};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < len) {
                        if ((unsigned)qq[k] & kReadyBit) {
                            fa[k] = rb + ((unsigned)(R + 1) << 4);
                            fg[k] = 0xffffffffu;
                        } else {
                            fa[k] = rb + (((unsigned)qq[k] & (unsigned)P.ring_mask) << 4);
                            fg[k] = ((unsigned)qq[k] & 0x3fffffffu) + 1u;
                        }
                        fc[k] = cc[k];
                    } else {
                        fa[k] = rb + ((unsigned)R << 4);
                        fg[k] = 0xffffffffu;
                        fc[k] = 0.0;
                    }
                    asm volatile("" : "+r"(fa[k]), "+r"(fg[k]));  // pin the values in registers before the gate
                }
            }
            const uint32_t myslot_s = smem_u32(myslot);
            const bool fast = (kind & 8) && w <= 4 && !remote_pending && P.fastpath;
            if (!wait_gate()) return;
            if (P.trace && lane == 0) P.trace[(size_t)c * 4 + 2] = clock64();
            int j = 0;
            if (fast) {
                bool okf;
                if (w <= 1) okf = fast_resolve<1, EXACT>(fa, fg, fc, acc, wd);
                else if (w == 2) okf = fast_resolve<2, EXACT>(fa, fg, fc, acc, wd);
                else if (w == 3) okf = fast_resolve<3, EXACT>(fa, fg, fc, acc, wd);
                else okf = fast_resolve<4, EXACT>(fa, fg, fc, acc, wd);
                if (!okf) return;
                xfin = P.unit_diag ? acc : (EXACT ? __ddiv_rn(acc, d) : acc * dinv);
                if (r >= 0) st_ring_s(myslot_s, xfin, mytag);
            } else {
            {
                double y0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
                const bool k0 = len > 0 && poll_dep(q0, P.mail, P.epoch, ring, P.ring_mask, y0);
                const bool k1 = len > 1 && poll_dep(q1, P.mail, P.epoch, ring, P.ring_mask, y1);
                const bool k2 = len > 2 && poll_dep(q2, P.mail, P.epoch, ring, P.ring_mask, y2);
                const bool k3 = len > 3 && poll_dep(q3, P.mail, P.epoch, ring, P.ring_mask, y3);
                if (k0) {
                    acc = sub_mul<EXACT>(acc, c0v, y0);
                    j = 1;
                    if (k1) {
                        acc = sub_mul<EXACT>(acc, c1v, y1);
                        j = 2;
                        if (k2) {
                            acc = sub_mul<EXACT>(acc, c2v, y2);
                            j = 3;
                            if (k3) {
                                acc = sub_mul<EXACT>(acc, c3v, y3);
                                j = 4;
                            }
                        }
                    }
                }
            }
            // ---- ordered hand-off loop.  Per lane: the dependency it is waiting for lives in registers as
            // (address, expected tag, coefficient).  The warp spins in a TIGHT loop (one 16-byte load, one
            // compare, one vote) until some lane's dependency has arrived; only then does it run the heavier
            // progress code for the lanes that can move.
            bool fin = false;
            bool is_glob = false;
            uint32_t saddr = 0;
            const ulonglong2* gaddr = P.mail;
            unsigned long long etag = 0ull;
            double v = 0.0;
            // walk forward from dependency j: consume everything already available, park on the first that is not
            auto advance = [&]() {
                while (j < len) {
                    const int e_ = sidx[j * 32 + lane];
                    const double v_ = sval[j * 32 + lane];
                    double x_;
                    if (poll_dep(e_, P.mail, P.epoch, ring, P.ring_mask, x_)) {
                        acc = sub_mul<EXACT>(acc, v_, x_);
                        ++j;
                    } else {
                        v = v_;
                        is_glob = e_ >= 0;
                        if (is_glob) {
                            gaddr = P.mail + e_;
                            etag = P.epoch;
                        } else {
                            const int lp = e_ & 0x3fffffff;
                            saddr = smem_u32(ring + (lp & P.ring_mask));
                            etag = (unsigned long long)lp + 1ull;
                        }
                        return;
                    }
                }
                xfin = P.unit_diag ? acc : (EXACT ? __ddiv_rn(acc, d) : acc * dinv);
                if (r >= 0) st_ring(myslot, xfin, mytag);
                fin = true;
            };
            advance();
            while (!__all_sync(FULL, fin)) {
                double xv = 0.0;
                bool ok = false;
                unsigned spins = 0;
                while (true) {  // tight wait
                    if (!fin) {
                        unsigned long long xb, t;
                        if (is_glob)
                            ld_b128_gpu(gaddr, xb, t);
                        else
                            ld_b128_cta(saddr, xb, t);
                        ok = (t == etag);
                        xv = __longlong_as_double((long long)xb);
                    }
                    if (__any_sync(FULL, ok)) break;
                    if ((++spins & 0xfffu) == 0 && wd.expired_now(103)) return;
                }
                if (ok) {
                    acc = sub_mul<EXACT>(acc, v, xv);
                    ++j;
                    advance();
                }
            }
            }  // generic path
        } else {
            // ------------------------------------------------------------ warp-cooperative chunk
            if (!wait_gate()) return;
            int off = len;  // exclusive scan of len
            for (int s = 1; s < 32; s <<= 1) {
                const int t = __shfl_up_sync(FULL, off, s);
                if (lane >= s) off += t;
            }
            off -= len;
            const size_t base = (size_t)cp * 32;
            for (int rr = 0; rr < 32; ++rr) {
                const int rrow = __shfl_sync(FULL, r, rr);
                if (rrow < 0) continue;
                const int rl = __shfl_sync(FULL, len, rr);
                const int ro = __shfl_sync(FULL, off, rr);
                double part = 0.0;
                double run = __shfl_sync(FULL, acc, rr);  // EXACT: running value b - sum, in stored order
                for (int tb = 0; tb < rl; tb += 32) {
                    const int t = tb + lane;
                    const bool active = t < rl;
                    const int e = active ? __ldg(P.idx + base + ro + t) : 0;
                    const double v = active ? __ldg(P.val + base + ro + t) : 0.0;
                    double xv = 0.0;
                    bool ok = !active;
                    while (true) {
                        if (!ok) ok = poll_dep(e, P.mail, P.epoch, ring, P.ring_mask, xv);
                        if (__all_sync(FULL, ok)) break;
                        if (wd.expired(104)) return;
                    }
                    if (EXACT) {
                        // the reference's sequential accumulation: one rounded product per entry, subtracted
                        // in stored order (ordered shuffle pass instead of a tree)
                        const double prod = active ? __dmul_rn(v, xv) : 0.0;
                        const int cnt = min(32, rl - tb);
                        for (int l = 0; l < cnt; ++l) run = __dsub_rn(run, __shfl_sync(FULL, prod, l));
                    } else if (active) {
                        part = fma(v, xv, part);
                    }
                }
                if (!EXACT)
                    for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(FULL, part, s);
                if (lane == rr) {
                    const double t = EXACT ? run : acc - part;
                    xfin = P.unit_diag ? t : (EXACT ? __ddiv_rn(t, d) : t / d);
                    st_ring(myslot, xfin, mytag);
                }
                __syncwarp();
            }
        }

        if (P.trace && lane == 0) P.trace[(size_t)c * 4 + 3] = clock64();
        // ---------------------------------------------------------------- progress word, then global memory
        // (the release fence of the progress word must not wait for the global stores below)
        __syncwarp();
        if (lane == 0) st_release_cta_s32(&warp_cur[warp], has_next ? nlc : INT_MAX);
        if (r >= 0) {
            P.x[r] = xfin;
            if ((pubm >> lane) & 1u) st_mail(P.mail + r, xfin, P.epoch);
        }
        if (has_next && P.nstages == 1) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) issue(nlc, it + 1);
        }
        // rotate the software pipeline
        r = r1;
        len = len1;
        d = d1;
        acc = r >= 0 ? __ldcg(P.x + r) : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// level-scheduled variant: thread per row of one level, straight from the CSR-read arrays
template <bool EXACT>
__global__ void k_trsv_level(SrcDev S, const double* __restrict__ lu, int unit_diag, const int* __restrict__ lvl_rows,
                             int begin, int end, double* x) {
    const int s = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= end) return;
    const int row = lvl_rows[s];
    const int b = src_beg(S, row), e = src_end(S, row);
    double acc = x[row];
    for (int j = 0; j < e - b; ++j) {
        const int q = S.reverse ? e - 1 - j : b + j;
        const int vp = S.val_pos ? S.val_pos[q] : q;
        acc = sub_mul<EXACT>(acc, lu[vp], __ldcg(x + S.dep_idx[q]));
    }
    if (!unit_diag) {
        const double d = lu[S.diag_pos[row]];
        acc = EXACT ? __ddiv_rn(acc, d) : acc / d;
    }
    x[row] = acc;
}

// ------------------------------------------------------------------------------------------------
// reference-layout factor extraction (wrap_matrix / factors_info, binding.cpp:118-175)
//   mode 0: unit lower  = strict lower of lu + explicit 1.0 diagonal LAST   (ILU0.hpp:82-98)
//   mode 1: upper       = diagonal FIRST + strict upper
//   mode 2: lower incl. diagonal (IChol0 L, diagonal last)
__global__ void k_factor_counts(int n, const int* __restrict__ ptr, const int* __restrict__ dpos, int mode,
                                int* __restrict__ outptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    outptr[i + 1] = mode == 1 ? ptr[i + 1] - dpos[i] : dpos[i] - ptr[i] + 1;
    if (i == 0) outptr[0] = 0;
}
__global__ void k_factor_fill(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                              const int* __restrict__ dpos, const double* __restrict__ lu, int mode,
                              const int* __restrict__ outptr, int* __restrict__ oidx, double* __restrict__ oval) {
    // 8 lanes per row
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t >> 3;
    const int gl = (int)(t & 7);
    if (i >= n) return;
    const int b = mode == 1 ? dpos[i] : ptr[i];
    const int e = mode == 1 ? ptr[i + 1] : dpos[i] + 1;
    const int o = outptr[i];
    for (int p = b + gl; p < e; p += 8) {
        oidx[o + p - b] = idx[p];
        oval[o + p - b] = (mode == 0 && p == dpos[i]) ? 1.0 : lu[p];
    }
}

inline int blocks_for(int64_t work, int threads) { return (int)((work + threads - 1) / threads); }

}  // namespace

void run_solve(ilub_handle* h, int slot, double* x, int64_t ldx, int nrhs, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    ILUB_REQUIRE(T.built, ILUB_ERR_INVALID, "solve not analysed");
    const Options& opt = options();
    const bool exact = opt.exact != 0;
    if (opt.solve_algo == 1) {
        build_level_sets(h, slot, s);
        std::vector<int> lptr(T.nlevels + 1);
        ILUB_CUDA(cudaMemcpyAsync(lptr.data(), T.lvl_ptr.p, sizeof(int) * (T.nlevels + 1), cudaMemcpyDeviceToHost, s));
        ILUB_CUDA(cudaStreamSynchronize(s));
        const SolveSource src = solve_source(h, slot);
        const SrcDev S = to_dev(src);
        for (int c = 0; c < nrhs; ++c) {
            double* xc = x + (int64_t)c * ldx;
            for (int l = 0; l < T.nlevels; ++l) {
                const int cnt = lptr[l + 1] - lptr[l];
                if (cnt <= 0) continue;
                if (exact)
                    k_trsv_level<true><<<blocks_for(cnt, 128), 128, 0, s>>>(S, h->lu.p, src.unit_diag ? 1 : 0,
                                                                              T.lvl_rows.p, lptr[l], lptr[l + 1], xc);
                else
                    k_trsv_level<false><<<blocks_for(cnt, 128), 128, 0, s>>>(S, h->lu.p, src.unit_diag ? 1 : 0,
                                                                               T.lvl_rows.p, lptr[l], lptr[l + 1], xc);
                h->launches++;
            }
        }
        ILUB_CUDA(cudaGetLastError());
        return;
    }

    TrsvParams P;
    P.part_chunk_ptr = T.part_chunk_ptr.p;
    P.chunk_ptr = T.chunk_ptr.p;
    P.ckind = T.ckind.p;
    P.rows = T.rows.p;
    P.rlen = T.rlen.p;
    P.idx = T.idx.p;
    P.val = T.val.p;
    P.diag = T.diag.p;
    P.pubmask = T.pubmask.p;
    P.gate = T.gate.p;
    P.slack = (int)opt.gate_slack;
    P.fastpath = (int)opt.fastpath;
    P.mail = T.mail.p;
    P.ring_mask = T.cfg_ring_chunks * 32 - 1;
    P.rc = T.cfg_ring_chunks;
    P.near = T.cfg_near;
    P.nstages = T.cfg_stages;
    P.stage_slabs = T.cfg_stage_slabs;
    P.unit_diag = T.unit_diag ? 1 : 0;
    P.status = h->status.p;
    P.spin_limit = (long long)opt.spin_limit_ms * h->clock_khz;
    if (opt.trace && T.trace.n < (size_t)T.nchunks * 4) T.trace.alloc((size_t)T.nchunks * 4);
    P.trace = opt.trace ? T.trace.p : nullptr;
    static thread_local bool attr_set[2] = {false, false};
    if (!attr_set[exact ? 1 : 0]) {
        if (exact)
            ILUB_CUDA(cudaFuncSetAttribute(k_trsv_syncfree<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024));
        else
            ILUB_CUDA(cudaFuncSetAttribute(k_trsv_syncfree<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024));
        attr_set[exact ? 1 : 0] = true;
    }
    for (int c = 0; c < nrhs; ++c) {
        P.epoch = ++T.epoch;  // 64-bit tag: never wraps, the mailbox is never reset
        P.x = x + (int64_t)c * ldx;
        if (exact)
            k_trsv_syncfree<true><<<T.nparts, T.cfg_warps * 32, T.cfg_smem, s>>>(P);
        else
            k_trsv_syncfree<false><<<T.nparts, T.cfg_warps * 32, T.cfg_smem, s>>>(P);
        h->launches++;
    }
    ILUB_CUDA(cudaGetLastError());
}

void extract_factor(ilub_handle* h, int mode, double* data, int* indices, int* indptr, cudaStream_t s) {
    const int n = h->n;
    k_factor_counts<<<blocks_for(n, 256), 256, 0, s>>>(n, h->ptr.p, h->dpos.p, mode, indptr);
    size_t tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, indptr, indptr, n + 1, s);
    DevBuf<unsigned char> tmp;
    tmp.alloc(tb);
    ILUB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, indptr, indptr, n + 1, s));
    k_factor_fill<<<blocks_for((int64_t)n * 8, 256), 256, 0, s>>>(n, h->ptr.p, h->idx.p, h->dpos.p, h->lu.p, mode,
                                                                  indptr, indices, data);
    h->launches += 3;
    ILUB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace ilub
