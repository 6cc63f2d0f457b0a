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

#include <cooperative_groups.h>

#include <cstring>

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
    const int* hptr;
    const int* halo_src;
    const int* halo_cons;
    int halo_mask;
    int nfetch;
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

// Shared-memory geometry of one CTA (all dependency polls are 16-byte shared-memory loads).
struct SmemView {
    uint32_t ring_s;   // shared address of the x ring  {x, local position + 1}
    uint32_t halo_s;   // shared address of the halo ring {x, edge number + 1}
    unsigned ring_mask, halo_mask;
    // (address, expected tag) of an encoded dependency
    __device__ __forceinline__ void locate(int e, uint32_t& addr, unsigned& tag) const {
        const unsigned p = (unsigned)e & 0x3fffffffu;
        const bool halo = ((unsigned)e & kHaloBit) != 0;
        addr = (halo ? halo_s : ring_s) + ((p & (halo ? halo_mask : ring_mask)) << 4);
        tag = p + 1u;
    }
};

__device__ __forceinline__ bool poll_dep(int e, const SmemView& sv, const ulonglong2* mail, unsigned long long epoch,
                                         double& xv) {
    if (e >= 0) return ld_mail(mail + e, epoch, xv);  // direct mailbox poll (halo-heavy chunks only)
    uint32_t a;
    unsigned g;
    sv.locate(e, a, g);
    unsigned long long xb, t;
    ld_b128_cta(a, xb, t);
    xv = __longlong_as_double((long long)xb);
    return (unsigned)t == g;
}

// Critical path of a single-level lane-per-row chunk: NW ring polls in flight per lane (absent dependencies
// point at an always-ready dummy entry holding 0.0), one vote, NW multiply-subtracts in stored order.
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
    const int WC = W - P.nfetch;  // compute warps; the last nfetch warps are halo fetchers
    const int R = P.ring_mask + 1;
    const int HR = P.halo_mask + 1;
    ulonglong2* ring = reinterpret_cast<ulonglong2*>(smem);                  // [R + 8]; [R] = {0.0, ~0} dummy
    ulonglong2* halo = ring + (R + 8);                                       // [HR]
    unsigned char* stage_base = reinterpret_cast<unsigned char*>(halo + HR);
    const int stage_bytes = P.stage_slabs * 384;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(stage_base + (size_t)WC * P.nstages * stage_bytes);
    uint64_t* cbar = mbar + WC * P.nstages;  // [rc] one count-1 barrier per ring chunk slot: phase k = chunk slot + k*rc
    int* warp_cur = reinterpret_cast<int*>(cbar + P.rc);

    const int c0 = P.part_chunk_ptr[blockIdx.x];
    const int nloc = P.part_chunk_ptr[blockIdx.x + 1] - c0;
    for (int i = threadIdx.x; i < R + 8 + HR; i += blockDim.x) ring[i] = make_ulonglong2(0ull, 0ull);
    if (threadIdx.x < 32) warp_cur[threadIdx.x] = (threadIdx.x < WC && threadIdx.x < nloc) ? threadIdx.x : INT_MAX;
    if (lane == 0 && warp < WC)
        for (int s = 0; s < P.nstages; ++s) mbar_init(&mbar[warp * P.nstages + s], 1);
    for (int i = threadIdx.x; i < P.rc; i += blockDim.x) mbar_init(&cbar[i], 1);
    fence_mbar_init();
    __syncthreads();
    if (threadIdx.x == 0) ring[R] = make_ulonglong2(0ull, 0xffffffffffffffffull);
    __syncthreads();

    Watchdog wd(P.status, P.spin_limit);
    SmemView sv;
    sv.ring_s = smem_u32(ring);
    sv.halo_s = smem_u32(halo);
    sv.ring_mask = (unsigned)P.ring_mask;
    sv.halo_mask = (unsigned)P.halo_mask;
    int done_cache = 0;
    auto refresh_done = [&]() {
        int v = lane < WC ? ld_relaxed_cta_s32(&warp_cur[lane]) : INT_MAX;
        done_cache = __reduce_min_sync(FULL, v);
    };

    if (warp >= WC) {
        // ================================================================== halo fetcher warps
        // Remote edges (dependencies on rows of other CTAs, or too far back for the x ring) are numbered per
        // CTA in the order their consumers are processed.  A fetcher takes 128 consecutive edges, polls the
        // producers' 16-byte mailboxes in global memory {x, epoch} and drops each value into the halo ring
        // {x, edge + 1}, far ahead of the compute warps -- which therefore never wait on an L2 round trip.
        const int h0 = P.hptr[c0];
        const int nh = P.hptr[c0 + nloc] - h0;
        const int fid = warp - WC;
        for (int pos = fid * 128; pos < nh; pos += P.nfetch * 128) {
            int src[4], need = 0;
            bool todo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int hp = pos + k * 32 + lane;
                todo[k] = hp < nh;
                src[k] = todo[k] ? P.halo_src[h0 + hp] : 0;
                // the slot still belongs to edge hp - HR until that edge's consumer chunk has completed
                if (todo[k] && hp >= HR) need = max(need, P.halo_cons[h0 + hp - HR] + 1);
            }
            need = __reduce_max_sync(FULL, need);
            while (done_cache < need) {
                refresh_done();
                if (done_cache >= need) break;
                __nanosleep(200);
                if (wd.expired(108)) return;
            }
            unsigned spins = 0;
            while (true) {
                // all four mailbox loads of the lane are in flight together: one L2 round trip per poll round
                unsigned long long xb[4], tg[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (todo[k]) ld_b128_gpu(P.mail + src[k], xb[k], tg[k]);
                bool pending = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (todo[k]) {
                        if (tg[k] == P.epoch) {
                            const unsigned hp = (unsigned)(pos + k * 32 + lane);
                            st_ring_s(sv.halo_s + ((hp & sv.halo_mask) << 4), __longlong_as_double((long long)xb[k]),
                                      (unsigned long long)hp + 1ull);
                            todo[k] = false;
                        } else {
                            pending = true;
                        }
                    }
                }
                if (!__any_sync(FULL, pending)) break;
                if ((++spins & 0xffu) == 0 && wd.expired_now(109)) return;
            }
        }
        return;
    }

    // ====================================================================== compute warps
    if (warp >= nloc) return;
    unsigned phasebits = 0;

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

    for (int it = 0; lc < nloc; lc += WC, ++it) {
        const int c = c0 + lc;
        const int nlc = lc + WC;
        const bool has_next = nlc < nloc;
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
        const int gate_lp = P.gate[c];
        double xfin = 0.0;

        // (1) ring safety: the ring slots of chunk lc are those of chunk lc - rc; they may be overwritten once
        //     every chunk that could still read them through shared memory (ids <= lc - rc + near) is done.
        const int need = lc - P.rc + P.near + 1;
        while (need > done_cache) {
            refresh_done();
            if (done_cache >= need) break;
            __nanosleep(64);
            if (wd.expired(105)) return;
        }
        // (2) gate row: the last row (in processing order) that feeds this chunk's lowest level.  Until it is
        //     published nothing in this chunk can finish, so the warp waits on that ONE ring entry (one lane
        //     polls: one shared-memory wavefront per round) instead of 32 lanes polling 32 entries and
        //     saturating the shared-memory pipe for the warps on the critical path.  The gate row has a
        //     strictly lower level than every row of this chunk, so blocking on it cannot create a cycle.
        auto wait_gate = [&]() -> bool {
            if (gate_lp < 0 || P.slack < 0) return true;
            // Hardware-suspended wait on the completion barrier of the gate chunk: no polling traffic from the
            // (many) warps that are still far from the frontier, prompt wake-up when the chunk's warp arrives.
            // Parity is unambiguous: ring safety (1) guarantees chunk gate - rc has completed, and chunk
            // gate + rc cannot complete before this chunk does.
            const int gate_chunk = gate_lp >> 5;
            uint64_t* gb = &cbar[gate_chunk & (P.rc - 1)];
            const unsigned gpar = (unsigned)(gate_chunk / P.rc) & 1u;
            unsigned spins = 0;
            while (!mbar_try_wait(gb, gpar))
                if ((++spins & 0xffu) == 0 && wd.expired_now(106)) return false;
            return true;
        };
        if (P.trace && lane == 0) P.trace[(size_t)c * 4 + 1] = clock64();
        const uint32_t myslot_s = sv.ring_s + (((unsigned)(lc * 32 + lane) & sv.ring_mask) << 4);
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
            // everything the fast path needs is computed BEFORE the gate: what runs after it is the critical path
            uint32_t fa[4];
            unsigned fg[4];
            double fc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < len && !(kind & 16)) {
                    sv.locate(sidx[k * 32 + lane], fa[k], fg[k]);
                    fc[k] = sval[k * 32 + lane];
                } else {
                    fa[k] = sv.ring_s + ((unsigned)R << 4);  // dummy entry {0.0, ~0}
                    fg[k] = 0xffffffffu;
                    fc[k] = 0.0;
                }
                asm volatile("" : "+r"(fa[k]), "+r"(fg[k]));  // pin the values in registers before the gate
            }
            const bool fast = (kind & 8) && !(kind & 16) && w <= 4 && P.fastpath;
            if (!wait_gate()) return;
            if (P.trace && lane == 0) P.trace[(size_t)c * 4 + 2] = clock64();
            if (fast) {
                // ---- straight-line fast path: a single-level chunk has no internal dependencies and every
                // dependency sits at a lower level, so the whole warp may block until all of them have arrived.
                bool okf;
                if (w <= 1) okf = fast_resolve<1, EXACT>(fa, fg, fc, acc, wd);
                else if (w == 2) okf = fast_resolve<2, EXACT>(fa, fg, fc, acc, wd);
                else if (w == 3) okf = fast_resolve<3, EXACT>(fa, fg, fc, acc, wd);
                else okf = fast_resolve<4, EXACT>(fa, fg, fc, acc, wd);
                if (!okf) return;
                xfin = P.unit_diag ? acc : (EXACT ? __ddiv_rn(acc, d) : acc * dinv);
                if (r >= 0) st_ring_s(myslot_s, xfin, mytag);
            } else {
                // ---- general ordered hand-off loop (chunks that mix levels, rows with more than 4 entries).
                // Per lane the dependency it is waiting for lives in registers as (address, tag, coefficient);
                // the warp spins in a tight loop (one 16-byte load, one compare, one vote) until some lane's
                // dependency has arrived and only then runs the progress code for the lanes that can move.
                int j = 0;
                bool fin = false;
                uint32_t saddr = sv.ring_s + ((unsigned)R << 4);
                unsigned etag = 0u;
                const ulonglong2* gaddr = nullptr;  // non-null: parked on a direct mailbox dependency
                double v = 0.0;
                auto advance = [&]() {
                    while (j < len) {
                        const int e_ = sidx[j * 32 + lane];
                        const double v_ = sval[j * 32 + lane];
                        double x_;
                        if (poll_dep(e_, sv, P.mail, P.epoch, x_)) {
                            acc = sub_mul<EXACT>(acc, v_, x_);
                            ++j;
                        } else {
                            v = v_;
                            gaddr = e_ >= 0 ? P.mail + e_ : nullptr;
                            if (e_ < 0) sv.locate(e_, saddr, etag);
                            return;
                        }
                    }
                    xfin = P.unit_diag ? acc : (EXACT ? __ddiv_rn(acc, d) : acc * dinv);
                    if (r >= 0) st_ring_s(myslot_s, xfin, mytag);
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
                            if (gaddr) {
                                ld_b128_gpu(gaddr, xb, t);
                                ok = (t == P.epoch);
                            } else {
                                ld_b128_cta(saddr, xb, t);
                                ok = ((unsigned)t == etag);
                            }
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
            }
        } else {
            // ------------------------------------------------------------ warp-cooperative chunk
            if (!wait_gate()) return;
            if (P.trace && lane == 0) P.trace[(size_t)c * 4 + 2] = clock64();
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
                        if (!ok) ok = poll_dep(e, sv, P.mail, P.epoch, xv);
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
                    st_ring_s(myslot_s, xfin, mytag);
                }
                __syncwarp();
            }
        }

        if (P.trace && lane == 0) P.trace[(size_t)c * 4 + 3] = clock64();
        // ---------------------------------------------------------------- progress word, then global memory
        // (the release fence of the progress word must not wait for the global stores below)
        __syncwarp();
        if (lane == 0) mbar_arrive(&cbar[lc & (P.rc - 1)]);  // chunk complete: wakes the warps gated on it
        if (r >= 0 && ((pubm >> lane) & 1u)) st_mail(P.mail + r, xfin, P.epoch);  // remote consumers next
        if (lane == 0) st_release_cta_s32(&warp_cur[warp], has_next ? nlc : INT_MAX);
        if (r >= 0) P.x[r] = xfin;
        if (has_next && P.nstages == 1) {
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
    // the rows of a level lie anywhere in the matrix: every access below is a cache miss of its own, so the operands
    // of four entries are requested together (index + coefficient, then the four solution values) and consumed in the
    // reference's order
    double d = 1.0;
    if (!unit_diag) d = lu[S.diag_pos[row]];
    for (int j0 = 0; j0 < e - b; j0 += 4) {
        int dep[4];
        double c[4], xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dep[u] = -1;
            c[u] = 0.0;
            if (j0 + u < e - b) {
                const int q = S.reverse ? e - 1 - (j0 + u) : b + j0 + u;
                dep[u] = S.dep_idx[q];
                c[u] = lu[S.val_pos ? S.val_pos[q] : q];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = dep[u] >= 0 ? __ldcg(x + dep[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (dep[u] >= 0) acc = sub_mul<EXACT>(acc, c[u], xv[u]);
    }
    if (!unit_diag) acc = EXACT ? __ddiv_rn(acc, d) : acc / d;
    x[row] = acc;
}

// Level loop for general matrices with many, moderately wide levels (option solve_algo = 4; automatic for e.g. the
// factors of threshold factorisations: ~10^3 levels of a few hundred rows): ONE cooperative launch walks all levels
// straight from the CSR-read arrays, G lanes per row (EXACT: one lane per row, the reference's accumulation order),
// and every CTA passes a grid-wide barrier (monotone arrival counter, relaxed polls) between two levels.  The grid is
// sized to the average level -- the barrier costs one atomic and one polling thread per CTA.
template <bool EXACT, int G, bool CLUSTER>
__global__ void __launch_bounds__(512) k_trsv_levels(SrcDev S, const double* __restrict__ lu, int unit_diag,
                                                     const int* __restrict__ lvl_ptr, const int* __restrict__ lvl_rows,
                                                     int nlevels, double* x, unsigned* barrier, long long spin_limit,
                                                     int* status) {
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
    const int ng = blockDim.x / G;
    const int g = blockIdx.x * ng + threadIdx.x / G;
    const int gstride = gridDim.x * ng;
    long long t0 = 0;
    // Everything about a row except the values it depends on is static: the first row a lane group takes in level
    // l + 1 (its extent, right-hand side, diagonal and up to kPre entries per lane) is fetched BEFORE the barrier of
    // level l, so that after the barrier only the gather of x, the reduction and the store remain.
    constexpr int kPre = 3;
    int prow = -1, pb = 0, pe = 0, pidx[kPre];
    double pval[kPre], prhs = 0.0, pdiag = 1.0;
    auto prefetch = [&](int l) {
        prow = -1;
        if (EXACT || l >= nlevels) return;
        const int sidx = __ldg(lvl_ptr + l) + g;
        if (sidx >= __ldg(lvl_ptr + l + 1)) return;
        prow = lvl_rows[sidx];
        pb = src_beg(S, prow);
        pe = src_end(S, prow);
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int j = gl + k * G;
            pidx[k] = -1;
            pval[k] = 0.0;
            if (j < pe - pb) {
                const int q = S.reverse ? pe - 1 - j : pb + j;
                pidx[k] = S.dep_idx[q];
                pval[k] = lu[S.val_pos ? S.val_pos[q] : q];
            }
        }
        if (gl == 0) {
            prhs = __ldcg(x + prow);
            pdiag = unit_diag ? 1.0 : lu[S.diag_pos[prow]];
        }
    };
    prefetch(0);
    for (int l = 0; l < nlevels; ++l) {
        const int lb = __ldg(lvl_ptr + l), le = __ldg(lvl_ptr + l + 1);
        int first = lb + g;
        if (!EXACT && prow >= 0) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < kPre; ++k)
                if (pidx[k] >= 0) acc = fma(pval[k], __ldcg(x + pidx[k]), acc);
            for (int j = gl + kPre * G; j < pe - pb; j += G) {
                const int q = S.reverse ? pe - 1 - j : pb + j;
                const int vp = S.val_pos ? S.val_pos[q] : q;
                acc = fma(lu[vp], __ldcg(x + S.dep_idx[q]), acc);
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
            if (gl == 0) x[prow] = unit_diag ? prhs - acc : (prhs - acc) / pdiag;
            first += gstride;
        }
        for (int sidx = first; sidx < le; sidx += gstride) {
            const int row = lvl_rows[sidx];
            const int b = src_beg(S, row), e = src_end(S, row);
            if (EXACT) {
                // the reference's accumulation order, one lane per row; the operands of eight entries are requested
                // together (independent loads), then consumed in order: same bits, an eighth of the round trips
                double acc = __ldcg(x + row);
                for (int j0 = 0; j0 < e - b; j0 += 8) {
                    double v8[8], x8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        v8[u] = 0.0;
                        x8[u] = 0.0;
                        if (j0 + u < e - b) {
                            const int q = S.reverse ? e - 1 - (j0 + u) : b + j0 + u;
                            v8[u] = lu[S.val_pos ? S.val_pos[q] : q];
                            x8[u] = __ldcg(x + S.dep_idx[q]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (j0 + u < e - b) acc = sub_mul<true>(acc, v8[u], x8[u]);
                }
                if (!unit_diag) acc = __ddiv_rn(acc, lu[S.diag_pos[row]]);
                x[row] = acc;
            } else {
                double acc = 0.0;
                for (int j = gl; j < e - b; j += G) {
                    const int q = S.reverse ? e - 1 - j : b + j;
                    const int vp = S.val_pos ? S.val_pos[q] : q;
                    acc = fma(lu[vp], __ldcg(x + S.dep_idx[q]), acc);
                }
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
                if (gl == 0) {
                    double r = __ldcg(x + row) - acc;
                    if (!unit_diag) r = r / lu[S.diag_pos[row]];
                    x[row] = r;
                }
            }
        }
        prefetch(l + 1);
        if (l + 1 == nlevels) break;
        // barrier: the values of level l are visible (x is read through L2) before anyone starts level l + 1
        if (CLUSTER) {
            // the whole grid is ONE thread-block cluster: the hardware cluster barrier (arrive.release / wait.acquire at
            // cluster scope) replaces the atomic counter in global memory -- a fraction of its latency
            cooperative_groups::this_cluster().sync();
            continue;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(barrier, 1u);
            const unsigned target = (unsigned)(l + 1) * gridDim.x;
            unsigned spins = 0;
            while (ld_relaxed_gpu_u32(barrier) < target) {
                if ((++spins & 1023u) == 0) {
                    if (t0 == 0) t0 = clock64();
                    if (clock64() - t0 > spin_limit || *(volatile int*)status != 0) {
                        atomicCAS(status, 0, 103);
                        break;
                    }
                }
            }
            __threadfence();
        }
        __syncthreads();
    }
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

// one level of the level-packed layout (few very wide levels): thread s resolves position s of the (level, row)
// listing; index / coefficient / diagonal streams are contiguous across the threads of a warp, only the solution vector
// is gathered (it fits the L2)
template <bool EXACT>
__global__ void k_trsv_level_packed(const int* __restrict__ rows, const int* __restrict__ ptr, const int* __restrict__ idx,
                                    const double* __restrict__ val, const double* __restrict__ diag, int begin, int end,
                                    double* x) {
    // Programmatic dependent launch: the kernel of the NEXT level may be scheduled right away; like this one it first
    // fetches everything that does not depend on the solve (its rows, extents, diagonal and the first four index /
    // coefficient pairs) and only then waits for the previous level's grid to complete -- launch latency and two of the
    // three dependent memory round trips of a level overlap with the level before.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int s = begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s < end;
    int row = 0, b = 0, e = 0;
    double d = 1.0;
    int dep[4] = {-1, -1, -1, -1};
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    if (live) {
        row = rows[s];
        b = ptr[s];
        e = ptr[s + 1];
        d = diag[s];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (b + u < e) {
                dep[u] = idx[b + u];
                c[u] = val[b + u];
            }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!live) return;
    double acc = x[row];
    for (int j0 = b; j0 < e; j0 += 4) {
        double xv[4];
        if (j0 > b) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                dep[u] = -1;
                c[u] = 0.0;
                if (j0 + u < e) {
                    dep[u] = idx[j0 + u];
                    c[u] = val[j0 + u];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = dep[u] >= 0 ? __ldcg(x + dep[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (dep[u] >= 0) acc = sub_mul<EXACT>(acc, c[u], xv[u]);
    }
    x[row] = EXACT ? __ddiv_rn(acc, d) : acc / d;
}

// many levels of moderate width: the cooperative level loop (the chunked kernel's contiguous row ranges serialise on
// such DAGs: measured 92 ms against 5 ms on ILUT factors of a 64^3 Poisson matrix)
bool levels_wanted(const ilub_handle* h, int slot) {
    const TriSolve& T = h->solve[slot];
    const int64_t algo = options().solve_algo;
    if (algo == 4) return true;
    if (algo != 0 || T.sys.built || deep_wanted(h, slot)) return false;
    return T.nlevels > 32 && (int64_t)h->n < (int64_t)T.nlevels * 16384;
}

// a handful of very wide levels (e.g. a randomly relabelled grid: 21 levels of 10^5 rows): the dependency wave is no
// problem, the rows of a level are scattered -- one plain launch per level straight from the CSR-read arrays beats the
// chunked kernel, whose contiguous row ranges turn every dependency into a remote one (measured on a randomly permuted
// 128^3 Poisson factor: 0.57 ms against 1.79 ms per apply; the cooperative level loop takes 0.79 ms)
bool few_levels_wanted(const ilub_handle* h, int slot) {
    const TriSolve& T = h->solve[slot];
    const int64_t algo = options().solve_algo;
    if (algo != 0 || T.sys.built || deep_wanted(h, slot)) return false;
    return T.nlevels <= 64 && (int64_t)h->n >= (int64_t)T.nlevels * 16384;
}

void run_solve(ilub_handle* h, int slot, double* x, int64_t ldx, int nrhs, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    ILUB_REQUIRE(T.built, ILUB_ERR_INVALID, "solve not analysed");
    const Options& opt = options();
    const bool exact = opt.exact != 0;
    if (opt.solve_algo == 0 && few_levels_wanted(h, slot)) {
        if (!T.lev.built) {
            build_level_packed(h, slot, s);
            pack_solve_values(h, slot, s);
        }
        const std::vector<int>& lptr = T.lvl_ptr_host;
        for (int c = 0; c < nrhs; ++c) {
            double* xc = x + (int64_t)c * ldx;
            for (int l = 0; l < T.nlevels; ++l) {
                const int cnt = lptr[l + 1] - lptr[l];
                if (cnt <= 0) continue;
                cudaLaunchConfig_t cfg;
                memset(&cfg, 0, sizeof(cfg));
                cfg.gridDim = dim3(blocks_for(cnt, 128), 1, 1);
                cfg.blockDim = dim3(128, 1, 1);
                cfg.stream = s;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                ILUB_CUDA(cudaLaunchKernelEx(&cfg, exact ? k_trsv_level_packed<true> : k_trsv_level_packed<false>,
                                             (const int*)T.lvl_rows.p, (const int*)T.lev.ptr.p, (const int*)T.lev.idx.p,
                                             (const double*)T.lev.val.p, (const double*)T.lev.diag.p, lptr[l], lptr[l + 1], xc));
                h->launches++;
            }
        }
        ILUB_CUDA(cudaGetLastError());
        return;
    }
    if (opt.solve_algo == 1) {
        build_level_sets(h, slot, s);
        const std::vector<int>& lptr = T.lvl_ptr_host;
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

    if (opt.solve_algo != 2 && opt.solve_algo != 3 && opt.solve_algo != 4 && T.sys.built) {
        run_systolic(h, slot, x, ldx, nrhs, s);
        return;
    }
    if (block_wanted(h, slot)) {
        if (!T.blk.built) analyse_block(h, slot, s);
        run_block(h, slot, x, ldx, nrhs, s);
        return;
    }
    if (deep_wanted(h, slot)) {
        if (!T.deep.built) {
            analyse_deep(h, slot, s);
            pack_deep(h, slot, s);
        }
        run_deep(h, slot, x, ldx, nrhs, s);
        return;
    }
    if (levels_wanted(h, slot)) {
        build_level_sets(h, slot, s);
        const SolveSource src = solve_source(h, slot);
        const SrcDev S = to_dev(src);
        const int unit = src.unit_diag ? 1 : 0;
        const int G = exact ? 1 : 8;
        const int64_t avg = ((int64_t)h->n + T.nlevels - 1) / T.nlevels;  // rows of an average level
        if (h->fbarrier.n < 1) h->fbarrier.alloc(1);
        const int* lvl_ptr = T.lvl_ptr.p;
        const int* lvl_rows = T.lvl_rows.p;
        const double* lu = h->lu.p;
        int nlevels = T.nlevels;
        unsigned* barrier = h->fbarrier.p;
        long long spin_limit = (long long)opt.spin_limit_ms * h->clock_khz * 40;
        int* status = h->status.p;
        int unit_arg = unit;
        SrcDev Sarg = S;
        // an average level that fits ONE pass of a thread-block cluster of 8 CTAs (4096 lanes) with room to spare: one
        // cluster, hardware barrier (64^3 ILUT factors, 297 rows per level: 3.9 ms against 4.9 ms); wider levels: a
        // cooperative grid sized to the average level, atomic-counter barrier (96^3, 664 rows per level: 8.5 ms against
        // 12.5 ms on the eight SMs of a cluster)
        const bool cluster = opt.levels_cluster != 0 && avg * G <= 2400;
        if (cluster) {
            typedef void (*kern_t)(SrcDev, const double*, int, const int*, const int*, int, double*, unsigned*, long long, int*);
            const kern_t kern = exact ? (kern_t)k_trsv_levels<true, 1, true> : (kern_t)k_trsv_levels<false, 8, true>;
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(8, 1, 1);
            cfg.blockDim = dim3(512, 1, 1);
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 8;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            for (int c = 0; c < nrhs; ++c) {
                double* xc = x + (int64_t)c * ldx;
                ILUB_CUDA(cudaLaunchKernelEx(&cfg, kern, Sarg, lu, unit_arg, lvl_ptr, lvl_rows, nlevels, xc, barrier, spin_limit, status));
                h->launches++;
            }
        } else {
            void* kern = exact ? (void*)k_trsv_levels<true, 1, false> : (void*)k_trsv_levels<false, 8, false>;
            const int block = 256;
            int occ = 0;
            ILUB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, 0));
            ILUB_REQUIRE(occ > 0, ILUB_ERR_CUDA, "level-loop solve kernel does not fit on an SM");
            int grid = (int)std::min<int64_t>((int64_t)occ * h->sm_count, std::max<int64_t>(4, (2 * avg * G + block - 1) / block));
            for (int c = 0; c < nrhs; ++c) {
                double* xc = x + (int64_t)c * ldx;
                ILUB_CUDA(cudaMemsetAsync(barrier, 0, sizeof(unsigned), s));
                void* args[] = {(void*)&Sarg, (void*)&lu, (void*)&unit_arg, (void*)&lvl_ptr, (void*)&lvl_rows, (void*)&nlevels,
                                (void*)&xc,   (void*)&barrier, (void*)&spin_limit, (void*)&status};
                ILUB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(block), args, 0, s));
                h->launches++;
            }
        }
        ILUB_CUDA(cudaGetLastError());
        return;
    }
    if (!T.generic_built) {
        analyse_generic(h, slot, s);
        pack_solve_values(h, slot, s);
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
    P.hptr = T.hptr.p;
    P.halo_src = T.halo_src.p;
    P.halo_cons = T.halo_cons.p;
    P.halo_mask = T.cfg_halo - 1;
    P.nfetch = T.cfg_fetch;
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
    static thread_local unsigned attr_set[2] = {0, 0};
    if (first_use_on_this_device(attr_set, exact ? 1 : 0)) {
        if (exact)
            ILUB_CUDA(cudaFuncSetAttribute(k_trsv_syncfree<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024));
        else
            ILUB_CUDA(cudaFuncSetAttribute(k_trsv_syncfree<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024));
    }
    {   // the sync-free kernel only makes progress when every CTA is resident (chunks wait for chunks of other CTAs):
        // checked against the occupancy of this very configuration; one launch per level needs no co-residency
        int occ = 0;
        if (exact)
            ILUB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trsv_syncfree<true>, T.cfg_warps * 32, T.cfg_smem));
        else
            ILUB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trsv_syncfree<false>, T.cfg_warps * 32, T.cfg_smem));
        if ((int64_t)occ * h->sm_count < T.nparts) {
            build_level_sets(h, slot, s);
            const std::vector<int>& lptr = T.lvl_ptr_host;
            const SolveSource src = solve_source(h, slot);
            const SrcDev S = to_dev(src);
            for (int c = 0; c < nrhs; ++c) {
                double* xc = x + (int64_t)c * ldx;
                for (int l = 0; l < T.nlevels; ++l) {
                    const int cnt = lptr[l + 1] - lptr[l];
                    if (cnt <= 0) continue;
                    if (exact)
                        k_trsv_level<true><<<blocks_for(cnt, 128), 128, 0, s>>>(S, h->lu.p, src.unit_diag ? 1 : 0, T.lvl_rows.p, lptr[l], lptr[l + 1], xc);
                    else
                        k_trsv_level<false><<<blocks_for(cnt, 128), 128, 0, s>>>(S, h->lu.p, src.unit_diag ? 1 : 0, T.lvl_rows.p, lptr[l], lptr[l + 1], xc);
                    h->launches++;
                }
            }
            ILUB_CUDA(cudaGetLastError());
            return;
        }
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
