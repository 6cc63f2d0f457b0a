// Symbolic phase on the GPU: structure validation + diagonal pointers, transposed patterns for the
// scatter-form solves, sync-free level-set analysis of each triangular DAG, row -> CTA partition,
// (part, level, row) ordering and the packed solve layout (32-row chunks, SELL slabs).
//
// Replaces what the reference does inline and sequentially: the diagonal search / nnzL,nnzU count
// of compute_ilu0 (ILU0.hpp:36-45), natural_triangular_part (sparse_implementation.h:2568-2592)
// and the implicit "rows in index order" schedule of triangular_solve (:4040-4087).  The level
// sets themselves have no reference counterpart.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cub/cub.cuh>

#include <climits>
#include <cstring>
#include <mutex>

#include "ilub_common.cuh"

namespace ilub {

namespace {
Options g_options;
std::mutex g_options_mu;
thread_local Options* tl_active_options = nullptr;
const OptionField kOptionFields[] = {
    {"factor_algo", &Options::factor_algo},
    {"solve_algo", &Options::solve_algo},
    {"exact", &Options::exact},
    {"ctas_per_sm", &Options::ctas_per_sm},
    {"warps_per_cta", &Options::warps_per_cta},
    {"near_window", &Options::near_window},
    {"ring_chunks", &Options::ring_chunks},
    {"wcap", &Options::wcap},
    {"stages", &Options::stages},
    {"spin_limit_ms", &Options::spin_limit_ms},
    {"partition", &Options::partition},
    {"nparts", &Options::nparts},
    {"trace", &Options::trace},
    {"fetch_warps", &Options::fetch_warps},
    {"halo_entries", &Options::halo_entries},
    {"fastpath", &Options::fastpath},
    {"level_align", &Options::level_align},
    {"align_min_rows", &Options::align_min_rows},
    {"gate_slack", &Options::gate_slack},
    {"part_kb", &Options::part_kb},
    {"factor_group", &Options::factor_group},
    {"sys_pz", &Options::sys_pz},
    {"sys_stages", &Options::sys_stages},
    {"sys_cluster", &Options::sys_cluster},
    {"sys_plane", &Options::sys_plane},
    {"zero_copy", &Options::zero_copy},
    {"sys_tile", &Options::sys_tile},
    {"tile_fetch", &Options::tile_fetch},
    {"tile_warps", &Options::tile_warps},
    {"tile_debug", &Options::tile_debug},
    {"tile_nst", &Options::tile_nst},
    {"tile_ring", &Options::tile_ring},
    {"tile_nap_fwd", &Options::tile_nap_fwd},
    {"tile_nap_fetch", &Options::tile_nap_fetch},
    {"tile_nap_tma", &Options::tile_nap_tma},
    {"plane_r", &Options::plane_r},
    {"plane_wpc", &Options::plane_wpc},
    {"plane_npc", &Options::plane_npc},
    {"levels_cluster", &Options::levels_cluster},
    {"sys_pdl", &Options::sys_pdl},
    {"deep_block", &Options::deep_block}};
}  // namespace

Options& options() { return tl_active_options ? *tl_active_options : g_options; }
const OptionField* option_fields(int* count) {
    if (count) *count = (int)(sizeof(kOptionFields) / sizeof(kOptionFields[0]));
    return kOptionFields;
}
int option_index(const char* key) {
    int n = 0;
    const OptionField* f = option_fields(&n);
    for (int i = 0; i < n; ++i)
        if (strcmp(f[i].name, key) == 0) return i;
    return -1;
}
OptionScope::OptionScope(ilub_handle* h) : prev(tl_active_options) {
    if (!h) return;
    {
        std::lock_guard<std::mutex> lock(g_options_mu);
        h->opt_effective = g_options;
    }
    for (const auto& ov : h->opt_overrides) h->opt_effective.*(kOptionFields[ov.first].member) = ov.second;
    tl_active_options = &h->opt_effective;
}
OptionScope::~OptionScope() { tl_active_options = prev; }
void set_default_option(int field, int64_t value) {
    std::lock_guard<std::mutex> lock(g_options_mu);
    g_options.*(kOptionFields[field].member) = value;
}

namespace {

inline int blocks_for(int64_t work, int threads) { return (int)((work + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------
// validation + diagonal pointer.  8 lanes per row stride over the row (coalesced), the lane that
// sees column == row records the position; a ballot over the group detects a missing diagonal.
// status[0] = 0 ok | 1 unsorted/duplicate | 2 out of range | 3 missing diagonal ; status[1] = row
__global__ void k_validate(int n, const int* __restrict__ ptr, const int* __restrict__ idx, int* __restrict__ dpos,
                           int* status) {
    const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
    const int64_t row64 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool valid = row64 < n;
    const int row = valid ? (int)row64 : 0;
    int err = 0;
    bool found = false;
    if (valid) {
        const int b = ptr[row], e = ptr[row + 1];
        if (e < b) err = 1;
        for (int p = b + gl; p < e; p += 8) {
            const int c = idx[p];
            if (c < 0 || c >= n) err = 2;
            if (p > b && idx[p - 1] >= c && !err) err = 1;
            if (c == row) {
                found = true;
                dpos[row] = p;
            }
        }
    }
    const unsigned fm = __ballot_sync(0xffffffffu, found);
    if (valid && !err && gl == 0 && ((fm >> (g * 8)) & 0xffu) == 0) err = 3;
    if (err) {
        if (atomicCAS(&status[0], 0, err) == 0) status[1] = row;
    }
}

__global__ void k_count_tri(int n, const int* __restrict__ ptr, const int* __restrict__ dpos,
                            unsigned long long* out /*[2] lower, upper*/) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long lo = 0, up = 0;
    if (i < n) {
        lo = (unsigned long long)(dpos[i] - ptr[i]);
        up = (unsigned long long)(ptr[i + 1] - dpos[i] - 1);
    }
    typedef cub::BlockReduce<unsigned long long, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    lo = BR(tmp).Sum(lo);
    __syncthreads();
    up = BR(tmp).Sum(up);
    if (threadIdx.x == 0) {
        atomicAdd(&out[0], lo);
        atomicAdd(&out[1], up);
    }
}

// ------------------------------------------------------------------------------------------------
// transposed pattern of the strict upper / strict lower triangle
__global__ void k_tri_count_cols(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                                 const int* __restrict__ dpos, int upper, int* __restrict__ colcnt,
                                 int* __restrict__ rowcnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = upper ? dpos[i] + 1 : ptr[i];
    const int e = upper ? ptr[i + 1] : dpos[i];
    rowcnt[i] = e - b;
    for (int p = b; p < e; ++p) atomicAdd(&colcnt[idx[p]], 1);
}

__global__ void k_tri_emit(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                           const int* __restrict__ dpos, int upper, const int* __restrict__ rowoff,
                           int* __restrict__ keys, unsigned long long* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = upper ? dpos[i] + 1 : ptr[i];
    const int e = upper ? ptr[i + 1] : dpos[i];
    int o = rowoff[i];
    for (int p = b; p < e; ++p, ++o) {
        keys[o] = idx[p];
        vals[o] = ((unsigned long long)(unsigned)i << 32) | (unsigned)p;
    }
}

__global__ void k_tri_unpack(int64_t nt, const unsigned long long* __restrict__ vals, int* __restrict__ tidx,
                             int* __restrict__ tpos) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    tidx[t] = (int)(vals[t] >> 32);
    tpos[t] = (int)(vals[t] & 0xffffffffu);
}

// ------------------------------------------------------------------------------------------------
// sync-free level analysis: one thread per row, rows handed out in dependency-safe order by a
// ticket counter (a warp only ever waits for rows whose ticket was drawn before its own, and those
// warps are running), levels published with volatile stores -- the value itself is the flag.
// Dependencies on rows of the SAME warp (the offset-1 chain of a lexicographic grid: 32 dependent hops through L2
// per warp if they went through memory) never touch memory: they are collected in a lane mask and resolved by one
// pass of 31 shuffles in lane order -- lane k is final when the pass reaches it, because its own in-warp
// dependencies are lanes below k.
__global__ void __launch_bounds__(256) k_levels(SrcDev S, int n, int* level, int* ticket, int* nlev,
                                                long long spin_limit, int* status, int nproc) {
    const int lane = threadIdx.x & 31;
    volatile int* vlevel = level;
    while (true) {
        int t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        const int64_t base = (int64_t)t * 32;
        if (base >= nproc) break;  // (nproc < n: only the first rows of the solve order -- a probe of the DAG's shape)
        const int64_t r = base + lane;
        const bool valid = r < nproc;
        const int row = valid ? (S.backward ? n - 1 - (int)r : (int)r) : 0;
        int p = valid ? src_beg(S, row) : 0;
        const int e = valid ? src_end(S, row) : 0;
        int lv = 0;
        unsigned inmask = 0;
        bool dead = false;
        const long long t0 = clock64();
        unsigned spins = 0;
        // first sweep: the levels of eight dependencies are requested together (independent loads: one L2 round trip
        // per eight entries, not one per entry); what was not ready yet -- in a deep DAG the few entries nearest to
        // the row -- is polled afterwards
        int lo = e, hi = -1;
        for (int q0 = p; q0 < e; q0 += 8) {
            int l8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                l8[u] = INT_MAX;
                if (q0 + u < e) {
                    const int dep = S.dep_idx[q0 + u];
                    const int64_t pd = S.backward ? (int64_t)n - 1 - dep : (int64_t)dep;  // solve position of the dependency
                    if (pd >= base && pd < r) inmask |= 1u << (int)(pd - base);
                    else l8[u] = vlevel[dep];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (l8[u] == INT_MAX) continue;
                if (l8[u] >= 0) {
                    lv = max(lv, l8[u] + 1);
                } else {
                    lo = min(lo, q0 + u);
                    hi = max(hi, q0 + u);
                }
            }
        }
        for (int q = lo; q <= hi && !dead; ++q) {
            const int dep = S.dep_idx[q];
            const int64_t pd = S.backward ? (int64_t)n - 1 - dep : (int64_t)dep;
            if (pd >= base && pd < r) continue;
            int l;
            while ((l = vlevel[dep]) < 0) {
                if ((++spins & 255u) == 0 && (clock64() - t0 > spin_limit || *(volatile int*)status != 0)) {
                    dead = true;
                    break;
                }
            }
            if (!dead) lv = max(lv, l + 1);
        }
        if (__any_sync(0xffffffffu, dead)) {
            if (lane == 0) atomicCAS(status, 0, 100);
            return;
        }
#pragma unroll
        for (int k = 0; k < 31; ++k) {
            const int v = __shfl_sync(0xffffffffu, lv, k);
            if ((inmask >> k) & 1u) lv = max(lv, v + 1);
        }
        if (valid) vlevel[row] = lv;
        const int m = __reduce_max_sync(0xffffffffu, valid ? lv + 1 : 0);
        if (lane == 0) atomicMax(nlev, m);
    }
}

// level-packed layout: entry count of every position of the (level, row) listing, then the entries themselves
__global__ void k_lev_count(int n, SrcDev S, const int* __restrict__ rows, int* __restrict__ cnt) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int row = rows[s];
    cnt[s] = src_end(S, row) - src_beg(S, row);
}
__global__ void k_lev_fill(int n, SrcDev S, int unit_diag, const int* __restrict__ rows, const int* __restrict__ ptr,
                           int* __restrict__ idx, int* __restrict__ srcpos, int* __restrict__ dpos) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int row = rows[s];
    const int b = src_beg(S, row), e = src_end(S, row);
    const int p = ptr[s];
    for (int j = 0; j < e - b; ++j) {
        const int q = S.reverse ? e - 1 - j : b + j;  // the order the reference accumulates in
        idx[p + j] = S.dep_idx[q];
        srcpos[p + j] = S.val_pos ? S.val_pos[q] : q;
    }
    dpos[s] = unit_diag ? -1 : S.diag_pos[row];
}
// diagonal positions of the systems 1 .. nsys-1 of a batch: those of system 0 shifted by whole systems
__global__ void k_dpos_replicate(int n, int n1, int nnz1, int* __restrict__ dpos) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + n1;
    if (r < n) dpos[r] = dpos[r % n1] + (int)(r / n1) * nnz1;
}
// levels of the systems 1 .. nsys-1 of a batch = levels of system 0 (block-diagonal matrix, shared pattern)
__global__ void k_levels_replicate(int n, int n1, int* __restrict__ level) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + n1;
    if (r < n) level[r] = level[r % n1];
}
// complete 5-/7-point stencil on a lexicographic nx x ny x nz grid: the level of (i, j, k) is i + j + k (mirrored for
// backward solves)
__global__ void k_levels_grid(int n, int n1, int nx, int ny, int nz, int backward, int* __restrict__ level) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int q = (int)(r % n1);
    const int i = q % nx, j = (q / nx) % ny, k = q / (nx * ny);
    level[r] = backward ? (nx - 1 - i) + (ny - 1 - j) + (nz - 1 - k) : i + j + k;
}

// ------------------------------------------------------------------------------------------------
// partition: contiguous row ranges balanced by work (entries + per-row overhead)
struct WeightOp {
    SrcDev S;
    __host__ __device__ __forceinline__ long long operator()(int i) const {
        return (long long)(src_end(S, i) - src_beg(S, i)) + 2;
    }
};

constexpr int kPartShift = 33;  // sort key = part << 33 | level << 1 | has-remote-dependency

struct PartPlan {
    int mode;    // 0 contiguous ranges, 2 strips (2-D), 3 pencils (3-D)
    int nparts;
    int sy, sz;  // detected strides: y = (i % sz) / sy, z = i / sz   (mode 2: sz unused, y = i / sy)
    int ny, nz, gy, gz;
};

__global__ void k_assign_parts(int n, const long long* __restrict__ wsum /*inclusive*/, SrcDev S, PartPlan pl,
                               int* __restrict__ part) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p;
    if (pl.mode == 0) {  // contiguous row ranges balanced by work (entries + per-row overhead)
        const long long total = wsum[n - 1];
        const long long w = (long long)(src_end(S, i) - src_beg(S, i)) + 2;
        const long long start = wsum[i] - w;
        p = (int)((__int128)start * pl.nparts / total);
    } else if (pl.mode == 2) {  // strips of grid lines
        const int y = i / pl.sy;
        p = (int)((long long)y * pl.gy / pl.ny);
    } else {  // pencils along the fastest axis: a (y, z) tile for every x
        const int z = i / pl.sz;
        const int y = (i - z * pl.sz) / pl.sy;
        const int ty = (int)((long long)y * pl.gy / pl.ny);
        const int tz = (int)((long long)z * pl.gz / pl.nz);
        p = tz * pl.gy + ty;
    }
    if (p >= pl.nparts) p = pl.nparts - 1;
    part[i] = p;
}

// rows that read another part's result go last inside their level so that chunks are (mostly) either
// purely local or "remote" -- a warp polling global memory does not slow down local hand-offs
__global__ void k_sort_keys(int n, SrcDev S, const int* __restrict__ part, const int* __restrict__ level,
                            unsigned long long* __restrict__ keys, int* __restrict__ rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = part[i];
    unsigned remote = 0;
    for (int q = src_beg(S, i); q < src_end(S, i); ++q)
        if (part[S.dep_idx[q]] != p) {
            remote = 1;
            break;
        }
    keys[i] = ((unsigned long long)(unsigned)p << kPartShift) | ((unsigned long long)(unsigned)level[i] << 1) | remote;
    rows[i] = i;
}

// 64 sample rows; the one with the most dependencies reports its (absolute) offsets: out[0] = count, out[1..8]
__global__ void k_sample_offsets(int n, SrcDev S, int* out) {
    __shared__ unsigned long long best;
    if (threadIdx.x == 0) best = 0;
    __syncthreads();
    const long long step = n / 64 > 0 ? n / 64 : 1;
    long long r = (long long)threadIdx.x * step + step / 2 + (long long)threadIdx.x * 37;
    const int row = (int)(r % n);
    const int cnt = src_end(S, row) - src_beg(S, row);
    atomicMax(&best, ((unsigned long long)(unsigned)cnt << 32) | (unsigned)(63 - threadIdx.x));
    __syncthreads();
    if ((unsigned)(63 - threadIdx.x) == (unsigned)(best & 0xffffffffu)) {
        out[0] = cnt;
        for (int k = 0; k < cnt && k < 8; ++k) {
            int d = S.dep_idx[src_beg(S, row) + k] - row;
            out[1 + k] = d < 0 ? -d : d;
        }
    }
}

// how many dependency entries sit at one of the (<= 3) candidate offsets?
__global__ void k_count_offsets(int n, SrcDev S, int no, int o0, int o1, int o2, unsigned long long* out /*[2]*/) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long hit = 0, tot = 0;
    if (i < n) {
        for (int q = src_beg(S, i); q < src_end(S, i); ++q) {
            int d = S.dep_idx[q] - i;
            if (d < 0) d = -d;
            ++tot;
            if (d == o0 || (no > 1 && d == o1) || (no > 2 && d == o2)) ++hit;
        }
    }
    typedef cub::BlockReduce<unsigned long long, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    hit = BR(tmp).Sum(hit);
    __syncthreads();
    tot = BR(tmp).Sum(tot);
    if (threadIdx.x == 0) {
        atomicAdd(&out[0], hit);
        atomicAdd(&out[1], tot);
    }
}

// rows of part p occupy sorted positions [rstart[p], rstart[p+1]); chunks [pcp[p], pcp[p+1])
__global__ void k_part_bounds(int n, int nparts, const unsigned long long* __restrict__ skeys, int* rstart,
                              int* pcp) {
    // single block
    for (int p = threadIdx.x; p <= nparts; p += blockDim.x) {
        int lo = 0, hi = n;  // first sorted position whose part >= p
        while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            if ((int)(skeys[mid] >> kPartShift) < p)
                lo = mid + 1;
            else
                hi = mid;
        }
        rstart[p] = lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int p = 0; p < nparts; ++p) {
            pcp[p] = acc;
            acc += (rstart[p + 1] - rstart[p] + 31) >> 5;
        }
        pcp[nparts] = acc;
    }
}

__global__ void k_place_rows(int n, const unsigned long long* __restrict__ skeys, const int* __restrict__ order,
                             const int* __restrict__ rstart, const int* __restrict__ pcp, int* __restrict__ rows,
                             int* __restrict__ rowpos) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int p = (int)(skeys[s] >> kPartShift);
    const int rank = s - rstart[p];
    const int pos = ((pcp[p] + (rank >> 5)) << 5) | (rank & 31);
    const int row = order[s];
    rows[pos] = row;
    rowpos[row] = pos;
}

// ---- chunk assignment.  A "group" is a maximal run of sorted rows that must start a new chunk: a new part
// always does; with level alignment every (part, level) pair does, so that a chunk never mixes levels and has
// no internal dependencies (the straight-line fast path of the solve kernel relies on that).
__global__ void k_group_start(int n, const unsigned long long* __restrict__ skeys, int align_levels,
                              int* __restrict__ gstart) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int sh = align_levels ? 1 : kPartShift;   // drop the "remote" bit / everything below the part
    const bool head = s == 0 || (skeys[s] >> sh) != (skeys[s - 1] >> sh);
    gstart[s] = head ? s : 0;
}
__global__ void k_new_chunk(int n, const int* __restrict__ gstart, int* __restrict__ newchunk) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    newchunk[s] = ((s - gstart[s]) & 31) == 0 ? 1 : 0;
}
// pcp[p] = chunk of the first row of part p (parts are contiguous in sorted order; empty parts collapse)
__global__ void k_part_chunks(int n, int nparts, const unsigned long long* __restrict__ skeys,
                              const int* __restrict__ chunk_incl, int* __restrict__ pcp) {
    for (int p = threadIdx.x; p <= nparts; p += blockDim.x) {
        int lo = 0, hi = n;  // first sorted position whose part >= p
        while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            if ((int)(skeys[mid] >> kPartShift) < p)
                lo = mid + 1;
            else
                hi = mid;
        }
        pcp[p] = lo < n ? chunk_incl[lo] - 1 : chunk_incl[n - 1];
    }
}
__global__ void k_place_rows2(int n, const int* __restrict__ order, const int* __restrict__ gstart,
                              const int* __restrict__ chunk_incl, int* __restrict__ rows, int* __restrict__ rowpos) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int pos = ((chunk_incl[s] - 1) << 5) | ((s - gstart[s]) & 31);
    const int row = order[s];
    rows[pos] = row;
    rowpos[row] = pos;
}

// one warp per chunk: lengths, width, kind, slab count
__global__ void k_chunk_shape(int nchunks, SrcDev S, const int* __restrict__ rows, int wcap,
                              int* __restrict__ rlen, int* __restrict__ cslabs, unsigned char* __restrict__ ckind,
                              int* __restrict__ maxw_a, int* __restrict__ any_b) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= nchunks) return;
    const int row = rows[c * 32 + lane];
    const int len = row >= 0 ? src_end(S, row) - src_beg(S, row) : 0;
    rlen[c * 32 + lane] = len;
    const int w = __reduce_max_sync(0xffffffffu, len);
    const int tot = __reduce_add_sync(0xffffffffu, len);
    if (lane == 0) {
        if (w <= wcap) {
            ckind[c] = 0;
            cslabs[c] = w;
            atomicMax(maxw_a, w);
        } else {
            ckind[c] = 1;
            cslabs[c] = (tot + 31) >> 5;
            *any_b = 1;
        }
    }
}

// Pack one chunk.  mode 0: only count the chunk's remote edges (nrem[c]); mode 1: write everything, with the
// halo position of every remote edge taken from hptr (exclusive scan of nrem).
// Dependency encoding (see ilub_common.cuh): local ring position, or halo ring position for remote edges --
// every remote (other CTA / too far back) edge gets its own halo entry, filled by the CTA's fetcher warps
// from the producer's mailbox, so compute warps only ever poll shared memory.
__global__ void k_chunk_fill(int mode, int nchunks, SrcDev S, const int* __restrict__ rows,
                             const int* __restrict__ rlen, const int* __restrict__ chunk_ptr,
                             const int* __restrict__ rowpos, const int* __restrict__ pcp, int nparts, int near,
                             unsigned char* __restrict__ ckind, int* __restrict__ idx, int* __restrict__ srcpos,
                             int* __restrict__ diagpos, unsigned char* __restrict__ pub,
                             const int* __restrict__ level, int* __restrict__ gate, int* __restrict__ nrem,
                             const int* __restrict__ hptr, int* __restrict__ halo_src,
                             int* __restrict__ halo_cons, int halo_cap) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= nchunks) return;
    // owning part: last p with pcp[p] <= c
    int lo = 0, hi = nparts;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pcp[mid] <= c)
            lo = mid;
        else
            hi = mid;
    }
    const int cfirst = pcp[lo];
    const int row = rows[c * 32 + lane];
    const int len = rlen[c * 32 + lane];
    const int kind = ckind[c] & 1;
    const int b = row >= 0 ? src_beg(S, row) : 0;
    const int e = b + len;
    // remote edges of this lane, and their first halo position inside the chunk
    int myrem = 0;
    for (int j = 0; j < len; ++j) {
        const int q = S.reverse ? e - 1 - j : b + j;
        const int dc = rowpos[S.dep_idx[q]] >> 5;
        if (!(dc >= cfirst && dc <= c && c - dc <= near)) ++myrem;
    }
    int pre = myrem;
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, pre, d);
        if (lane >= d) pre += t;
    }
    const int chunk_rem = __shfl_sync(0xffffffffu, pre, 31);
    pre -= myrem;
    // a chunk with more remote edges than a fraction of the halo ring polls the mailboxes directly instead
    // (its edges could not all be resident in the ring at once)
    const bool use_halo = chunk_rem <= halo_cap;
    if (mode == 0) {
        if (lane == 0) nrem[c] = use_halo ? chunk_rem : 0;
        return;
    }
    const int hfirst = hptr[cfirst];           // halo positions are numbered per CTA
    int hp = hptr[c] + pre;                    // global halo position of this lane's next remote edge
    const int64_t base = (int64_t)chunk_ptr[c] * 32;
    const int w = chunk_ptr[c + 1] - chunk_ptr[c];
    diagpos[c * 32 + lane] = (row >= 0 && S.diag_pos) ? S.diag_pos[row] : -1;
    int off = 0;
    if (kind) {  // exclusive scan of len over the warp
        int v = len;
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += t;
        }
        off = v - len;
        const int tot = __shfl_sync(0xffffffffu, v, 31);
        for (int64_t t = tot + lane; t < (int64_t)w * 32; t += 32) {  // tail padding
            idx[base + t] = 0;
            srcpos[base + t] = -1;
        }
    }
    bool intra = false;
    const int lv = row >= 0 ? level[row] : 0x7fffffff;
    const int lvmin = __reduce_min_sync(0xffffffffu, lv);
    const int lvmax = __reduce_max_sync(0xffffffffu, row >= 0 ? lv : -1);
    int g = -1;  // last local position (CTA numbering) feeding the rows of this chunk's lowest level
    for (int j = 0; j < (kind ? len : w); ++j) {
        const int64_t at = kind ? base + off + j : base + (int64_t)j * 32 + lane;
        if (j >= len) {
            idx[at] = 0;
            srcpos[at] = -1;
            continue;
        }
        const int q = S.reverse ? e - 1 - j : b + j;
        const int d = S.dep_idx[q];
        const int pd = rowpos[d];
        const int dc = pd >> 5;
        unsigned enc;
        if (dc >= cfirst && dc <= c && c - dc <= near) {
            if (dc == c) intra = true;
            if (lv == lvmin) g = max(g, pd - (cfirst << 5));
            enc = kLocalBit | (unsigned)(pd - (cfirst << 5));
        } else if (use_halo) {
            enc = kLocalBit | kHaloBit | (unsigned)(hp - hfirst);
            halo_src[hp] = d;
            halo_cons[hp] = c - cfirst;
            ++hp;
            pub[d] = 1;  // row d must publish through the mailbox
        } else {
            enc = (unsigned)d;  // direct mailbox poll by the consumer
            pub[d] = 1;
        }
        idx[at] = (int)enc;
        srcpos[at] = S.val_pos ? S.val_pos[q] : q;
    }
    const bool any_intra = __any_sync(0xffffffffu, intra);
    if (lane == 0)
        ckind[c] = (unsigned char)(kind | (any_intra ? 2 : 0) | (chunk_rem > 0 ? 4 : 0) | (lvmin == lvmax ? 8 : 0) |
                                   ((chunk_rem > 0 && !use_halo) ? 16 : 0));
    g = __reduce_max_sync(0xffffffffu, g + 1) - 1;
    if (lane == 0) gate[c] = g;
}

__global__ void k_pubmask(int nchunks, const int* __restrict__ rows, const unsigned char* __restrict__ pub,
                          unsigned* __restrict__ pubmask) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= nchunks) return;
    const int row = rows[c * 32 + lane];
    const unsigned m = __ballot_sync(0xffffffffu, row >= 0 && pub[row] != 0);
    if (lane == 0) pubmask[c] = m;
}

__global__ void k_pack_values(int64_t nent, const int* __restrict__ srcpos, const double* __restrict__ lu,
                              double* __restrict__ val) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nent) return;
    const int s = srcpos[t];
    val[t] = s >= 0 ? lu[s] : 0.0;
}
__global__ void k_pack_diag(int64_t nent, const int* __restrict__ diagpos, const double* __restrict__ lu,
                            double* __restrict__ diag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nent) return;
    const int s = diagpos[t];
    diag[t] = s >= 0 ? lu[s] : 1.0;
}

__global__ void k_iota(int n, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
__global__ void k_level_ptr(int n, int nlevels, const int* __restrict__ slev, int* __restrict__ lptr) {
    // lptr[l] = first sorted position with level >= l
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > nlevels) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (slev[mid] < l)
            lo = mid + 1;
        else
            hi = mid;
    }
    lptr[l] = lo;
}

__global__ void k_expand_batched(int32_t n1, int64_t nnz1, int32_t nsys, const int* __restrict__ ptr1,
                                 const int* __restrict__ idx1, int* __restrict__ ptr, int* __restrict__ idx) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nnz = nnz1 * nsys;
    if (t < nnz) {
        const int s = (int)(t / nnz1);
        idx[t] = idx1[t - (int64_t)s * nnz1] + s * n1;
    }
    const int64_t n = (int64_t)n1 * nsys;
    if (t <= n) {
        const int s = (int)(t / n1);
        ptr[t] = (int)(ptr1[t - (int64_t)s * n1] + (int64_t)s * nnz1);
    }
}

int bits_for(unsigned v) {
    int b = 1;
    while (b < 32 && (v >> b)) ++b;
    return b;
}

}  // namespace

// ================================================================================================
int read_status(ilub_handle* h, cudaStream_t s) {
    int st[2] = {0, 0};
    ILUB_CUDA(cudaMemcpyAsync(st, h->status.p, sizeof(st), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    if (st[0] == 0) return 0;
    if (st[0] >= 100) {
        ILUB_CUDA(cudaMemsetAsync(h->status.p, 0, sizeof(int) * 4, s));
        throw Error{ILUB_ERR_WATCHDOG, "sync-free kernel exceeded its spin limit (code " + std::to_string(st[0]) + ")"};
    }
    const char* what = st[0] == 1   ? "column indices are not strictly ascending (unsorted or duplicate entries)"
                       : st[0] == 2 ? "column index out of range"
                                    : "structurally missing diagonal entry";
    throw Error{ILUB_ERR_STRUCTURE, std::string(what) + " in row " + std::to_string(st[1])};
}

void validate_and_find_diag(ilub_handle* h, cudaStream_t s) {
    const int n = h->n;
    // a batch is the block-diagonal repetition of ONE pattern (expand_batched): system 0 is checked, its diagonal
    // positions are shifted into the other systems
    const int nsys = h->nsys, n1 = n / nsys;
    const int64_t nnz1 = h->nnz / nsys;
    h->dpos.alloc(n);
    h->status.alloc(8);
    ILUB_CUDA(cudaMemsetAsync(h->status.p, 0, sizeof(int) * 8, s));
    k_validate<<<blocks_for((int64_t)n1 * 8, 256), 256, 0, s>>>(n1, h->ptr.p, h->idx.p, h->dpos.p, h->status.p);
    h->launches++;
    read_status(h, s);
    if (nsys > 1) {
        k_dpos_replicate<<<blocks_for((int64_t)n - n1, 256), 256, 0, s>>>(n, n1, (int)nnz1, h->dpos.p);
        h->launches++;
    }
    DevBuf<unsigned long long> cnt;
    cnt.alloc(2);
    ILUB_CUDA(cudaMemsetAsync(cnt.p, 0, 16, s));
    k_count_tri<<<blocks_for(n1, 256), 256, 0, s>>>(n1, h->ptr.p, h->dpos.p, cnt.p);
    h->launches++;
    unsigned long long hc[2];
    ILUB_CUDA(cudaMemcpyAsync(hc, cnt.p, 16, cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    h->nnz_lower = (int64_t)hc[0] * nsys;
    h->nnz_upper = (int64_t)hc[1] * nsys;
}

void build_transposed(ilub_handle* h, bool upper, cudaStream_t s) {
    Transposed& T = upper ? h->t_upper : h->t_lower;
    if (T.built) return;
    const int n = h->n;
    const int64_t nt = upper ? h->nnz_upper : h->nnz_lower;
    T.nt = nt;
    T.tptr.alloc(n + 1);
    T.tidx.alloc(nt);
    T.tpos.alloc(nt);
    DevBuf<int> colcnt, rowcnt, rowoff, keys, keys2;
    DevBuf<unsigned long long> vals, vals2;
    colcnt.alloc(n + 1);
    rowcnt.alloc(n);
    rowoff.alloc(n);
    keys.alloc(nt);
    keys2.alloc(nt);
    vals.alloc(nt);
    vals2.alloc(nt);
    ILUB_CUDA(cudaMemsetAsync(colcnt.p, 0, sizeof(int) * (n + 1), s));
    k_tri_count_cols<<<blocks_for(n, 256), 256, 0, s>>>(n, h->ptr.p, h->idx.p, h->dpos.p, upper ? 1 : 0, colcnt.p,
                                                        rowcnt.p);
    size_t tb1 = 0, tb2 = 0, tb3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb1, colcnt.p, T.tptr.p, n + 1, s);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, rowcnt.p, rowoff.p, n, s);
    cub::DoubleBuffer<int> dk(keys.p, keys2.p);
    cub::DoubleBuffer<unsigned long long> dv(vals.p, vals2.p);
    const int kb = bits_for((unsigned)n);
    cub::DeviceRadixSort::SortPairs(nullptr, tb3, dk, dv, (int)nt, 0, kb, s);
    DevBuf<unsigned char> tmp;
    tmp.alloc(std::max(tb1, std::max(tb2, tb3)));
    size_t tb = tmp.n;
    ILUB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, colcnt.p, T.tptr.p, n + 1, s));
    tb = tmp.n;
    ILUB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, rowcnt.p, rowoff.p, n, s));
    k_tri_emit<<<blocks_for(n, 256), 256, 0, s>>>(n, h->ptr.p, h->idx.p, h->dpos.p, upper ? 1 : 0, rowoff.p, keys.p,
                                                  vals.p);
    if (nt > 0) {
        tb = tmp.n;
        ILUB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, (int)nt, 0, kb, s));
        k_tri_unpack<<<blocks_for(nt, 256), 256, 0, s>>>(nt, dv.Current(), T.tidx.p, T.tpos.p);
    }
    h->launches += 6;
    ILUB_CUDA(cudaStreamSynchronize(s));  // temporaries die here
    T.built = true;
}

SolveSource solve_source(ilub_handle* h, int slot) {
    SolveSource S;
    const bool ichol = h->kind == KIND_ICHOL0;
    switch (slot) {
        case 0:  // gather, forward, strict lower of the CSR-read arrays
            S.beg_a = h->ptr.p;
            S.end_a = h->dpos.p;
            S.dep_idx = h->idx.p;
            S.unit_diag = !ichol;
            S.diag_pos = ichol ? h->dpos.p : nullptr;
            break;
        case 1:  // gather, backward, strict upper
            S.beg_a = h->dpos.p;
            S.beg_off = 1;
            S.end_a = h->ptr.p + 1;
            S.dep_idx = h->idx.p;
            S.diag_pos = h->dpos.p;
            S.backward = true;
            break;
        case 2:  // transposed strict upper, forward (scatter-form U^T solve)
            S.beg_a = h->t_upper.tptr.p;
            S.end_a = h->t_upper.tptr.p + 1;
            S.dep_idx = h->t_upper.tidx.p;
            S.val_pos = h->t_upper.tpos.p;
            S.diag_pos = h->dpos.p;
            break;
        default:  // 3: transposed strict lower, backward, entries accumulated in descending order
            S.beg_a = h->t_lower.tptr.p;
            S.end_a = h->t_lower.tptr.p + 1;
            S.dep_idx = h->t_lower.tidx.p;
            S.val_pos = h->t_lower.tpos.p;
            S.unit_diag = !ichol;
            S.diag_pos = ichol ? h->dpos.p : nullptr;
            S.backward = true;
            S.reverse_entries = true;
            break;
    }
    return S;
}

static PartPlan plan_partition(ilub_handle* h, const SrcDev& S, int64_t nent, cudaStream_t s) {
    const Options& opt = options();
    const int n = h->n;
    PartPlan pl;
    memset(&pl, 0, sizeof(pl));
    const int64_t maxp = (int64_t)h->sm_count * std::max<int64_t>(1, opt.ctas_per_sm);
    const int64_t bytes = 12 * nent + 16 * (int64_t)n;
    int64_t target = opt.nparts > 0 ? opt.nparts : std::max<int64_t>(1, bytes / (std::max<int64_t>(1, opt.part_kb) * 1024));
    target = std::min(target, maxp);
    target = std::min<int64_t>(target, std::max<int64_t>(1, ((int64_t)n + 31) / 32));
    pl.mode = 0;
    pl.nparts = (int)target;
    if (opt.partition == 1 || target == 1) return pl;
    DevBuf<int> dout;
    DevBuf<unsigned long long> dcnt;
    dout.alloc(16);
    dcnt.alloc(2);
    ILUB_CUDA(cudaMemsetAsync(dout.p, 0, sizeof(int) * 16, s));
    ILUB_CUDA(cudaMemsetAsync(dcnt.p, 0, 16, s));
    k_sample_offsets<<<1, 64, 0, s>>>(n, S, dout.p);
    int ho[16];
    ILUB_CUDA(cudaMemcpyAsync(ho, dout.p, sizeof(ho), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    h->launches++;
    const int cnt = ho[0];
    if (cnt < 1 || cnt > 3) return pl;
    int off[3] = {0, 0, 0};
    for (int k = 0; k < cnt; ++k) off[k] = ho[1 + k];
    std::sort(off, off + cnt);
    for (int k = 1; k < cnt; ++k)
        if (off[k] == off[k - 1]) return pl;
    k_count_offsets<<<blocks_for(n, 256), 256, 0, s>>>(n, S, cnt, off[0], off[1], off[2], dcnt.p);
    unsigned long long hc[2];
    ILUB_CUDA(cudaMemcpyAsync(hc, dcnt.p, 16, cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    h->launches++;
    if (hc[1] == 0 || (double)hc[0] < 0.97 * (double)hc[1]) return pl;
    if (cnt == 1) {
        if (off[0] == 1) pl.nparts = 1;  // a single chain: no parallelism to partition
        return pl;
    }
    if (cnt == 2) {
        pl.mode = 2;
        pl.sy = off[1];
        pl.ny = (n + pl.sy - 1) / pl.sy;
        pl.gy = (int)std::min<int64_t>(target, pl.ny);
        pl.gz = 1;
        pl.nparts = pl.gy;
        return pl;
    }
    pl.mode = 3;
    pl.sy = off[1];
    pl.sz = off[2];
    pl.ny = (pl.sz + pl.sy - 1) / pl.sy;
    pl.nz = (n + pl.sz - 1) / pl.sz;
    int g = (int)std::floor(std::sqrt((double)target));
    pl.gy = std::max(1, std::min(g, pl.ny));
    pl.gz = (int)std::max<int64_t>(1, std::min<int64_t>(target / pl.gy, pl.nz));
    pl.gy = (int)std::max<int64_t>(1, std::min<int64_t>(target / pl.gz, pl.ny));
    pl.nparts = pl.gy * pl.gz;
    return pl;
}

void analyse_solve(ilub_handle* h, int slot, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    if (T.built) return;
    if (slot == 2 || slot == 3) {
        DebugTimer t("  transposed pattern", s);
        build_transposed(h, slot == 2, s);
    }
    const Options& opt = options();
    const int n = h->n;
    const SolveSource src = solve_source(h, slot);
    const SrcDev S = to_dev(src);
    T.n = n;
    T.backward = src.backward;
    T.unit_diag = src.unit_diag;

    // ---- structured grids: line-systolic layout (ilub_systolic.cu).  The chunked sync-free layout below is then
    //      only built for slot 0 (the factorisation kernels walk it) or on demand (option solve_algo = 2).
    bool sys_ok;
    {
        DebugTimer t("  systolic layout", s);
        sys_ok = opt.solve_algo != 2 && analyse_systolic(h, slot, s);
    }

    // ---- levels.  The systems of a batch share one pattern: the dependency levels of ONE system are computed and
    //      replicated.  A complete grid stencil (every in-grid neighbour present, checked by the entry count on top of
    //      the per-entry verification of the systolic analysis) has the closed form i + j + k.
    const int nsys = h->nsys, n1 = n / nsys;
    T.level.alloc(n);
    bool closed_form = false;
    if (sys_ok) {
        const SysSolve& Y = T.sys;
        const int64_t nx = Y.nx, ny = Y.ny, nz = Y.nz1;
        const int64_t full = (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1);
        const int64_t have = ((slot == 0 || slot == 3) ? h->nnz_lower : h->nnz_upper) / nsys;
        closed_form = have == full;
    }
    if (closed_form) {
        DebugTimer t("  levels (closed form)", s);
        const SysSolve& Y = T.sys;
        k_levels_grid<<<blocks_for(n, 256), 256, 0, s>>>(n, n1, Y.nx, Y.ny, Y.nz1, src.backward ? 1 : 0, T.level.p);
        h->launches++;
        T.nlevels = Y.nx + Y.ny + Y.nz1 - 2;
    } else {
        ILUB_CUDA(cudaMemsetAsync(T.level.p, 0xff, sizeof(int) * n1, s));
        DevBuf<int> ctr;
        ctr.alloc(4);
        ILUB_CUDA(cudaMemsetAsync(ctr.p, 0, sizeof(int) * 4, s));
        const long long spin = (long long)opt.spin_limit_ms * h->clock_khz;  // kHz * ms = cycles
        int occ = 0;
        ILUB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_levels, 256, 0));
        // Large matrices are probed first: the levels of the first 65,536 rows of the solve order tell a deep narrow DAG
        // (fewer than 64 rows per level) from a wide one.  On a deep one the sync-free propagation below costs ~2.7 us per
        // level (0.3 s per solve at n = 10^6); the block layout (pattern only) is built instead and the levels come
        // out of the block-sequential kernel in the (max, +1) semiring (ilub_block.cu), exactly the same numbers.
        constexpr int kProbe = 65536;
        bool via_block = false;
        if (!sys_ok && nsys == 1 && opt.deep_block != 0 && opt.solve_algo == 0 && n1 >= 4 * kProbe) {
            DebugTimer t("  levels (probe)", s);
            const int grid = std::max(1, std::min(occ * h->sm_count, kProbe / 32));
            k_levels<<<grid, 256, 0, s>>>(S, n1, T.level.p, ctr.p, ctr.p + 1, spin * 4, h->status.p, kProbe);
            h->launches++;
            read_status(h, s);
            int hp[2];
            ILUB_CUDA(cudaMemcpy(hp, ctr.p, sizeof(hp), cudaMemcpyDeviceToHost));
            if ((int64_t)hp[1] * 64 > kProbe) {
                DebugTimer t2("  levels (block-sequential)", s);
                analyse_block(h, slot, s);
                T.nlevels = levels_block(h, slot, s);
                read_status(h, s);
                via_block = true;
            } else {
                ILUB_CUDA(cudaMemsetAsync(T.level.p, 0xff, sizeof(int) * n1, s));
                ILUB_CUDA(cudaMemsetAsync(ctr.p, 0, sizeof(int) * 4, s));
            }
        }
        if (!via_block) {
            {
                DebugTimer t("  levels", s);
                const int grid = std::max(1, std::min(occ * h->sm_count, (n1 + 31) / 32));
                k_levels<<<grid, 256, 0, s>>>(S, n1, T.level.p, ctr.p, ctr.p + 1, spin * 4, h->status.p, n1);
                h->launches++;
                if (nsys > 1) {
                    k_levels_replicate<<<blocks_for((int64_t)n - n1, 256), 256, 0, s>>>(n, n1, T.level.p);
                    h->launches++;
                }
            }
            read_status(h, s);
            int hctr[2];
            ILUB_CUDA(cudaMemcpy(hctr, ctr.p, sizeof(hctr), cudaMemcpyDeviceToHost));
            T.nlevels = hctr[1];
        }
    }
    T.built = true;
    if (T.blk.built && !deep_wanted(h, slot)) {  // (the probe was wrong about the DAG: the layout is of no use)
        BlockSolve& B = T.blk;
        B.fptr.release(); B.fpos.release(); B.fsrc.release(); B.nptr.release(); B.npos.release(); B.nsrc.release();
        B.cr.release(); B.co.release(); B.dsrc.release(); B.fval.release(); B.nval.release(); B.rdiag.release(); B.bmeta.release();
        B.llev.release(); B.bdepth.release(); B.cptr.release(); B.cpos.release(); B.cval.release();
        B.built = B.packed = false;
    }
    if (!sys_ok && block_wanted(h, slot)) {
        DebugTimer t("  block-sequential layout", s);
        analyse_block(h, slot, s);
        T.part_mode = 8;
        T.nparts = 1;
    } else if (!sys_ok && deep_wanted(h, slot)) {
        DebugTimer t("  level-loop layout", s);
        analyse_deep(h, slot, s);
        T.part_mode = 5;
        T.nparts = 1;
    } else if (!sys_ok && few_levels_wanted(h, slot)) {
        DebugTimer t("  level sets", s);
        build_level_packed(h, slot, s);  // rows and entries in (level, row) order: every level is one contiguous stream
        T.part_mode = 7;
        T.nparts = 0;
    } else if (!sys_ok && levels_wanted(h, slot)) {
        DebugTimer t("  level sets", s);
        build_level_sets(h, slot, s);  // the cooperative level loop walks the CSR-read arrays: no packed layout
        T.part_mode = 6;
        T.nparts = 0;
    } else if (!sys_ok) {  // (slot 0: the sync-free factorisation builds it on demand, factor_prepare)
        DebugTimer t("  chunked layout", s);
        analyse_generic(h, slot, s);
    }
}

void analyse_generic(ilub_handle* h, int slot, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    if (T.generic_built) return;
    const Options& opt = options();
    const int n = h->n;
    const SolveSource src = solve_source(h, slot);
    const SrcDev S = to_dev(src);
    DevBuf<int> ctr;
    ctr.alloc(4);

    // ---- partition + (part, level, remote, row) order
    const int64_t nent_est = (slot == 0 || slot == 3) ? h->nnz_lower : h->nnz_upper;
    const PartPlan pl = plan_partition(h, S, nent_est, s);
    const int nparts = pl.nparts;
    T.nparts = nparts;
    T.part_mode = pl.mode;
    T.stride_y = pl.sy;
    T.stride_z = pl.sz;
    T.grid_y = pl.gy;
    T.grid_z = pl.gz;
    DevBuf<long long> wsum;
    wsum.alloc(n);
    DevBuf<unsigned long long> keys, keys2;
    DevBuf<int> rowsb, rows2, part;
    keys.alloc(n);
    keys2.alloc(n);
    rowsb.alloc(n);
    rows2.alloc(n);
    part.alloc(n);
    cub::CountingInputIterator<int> cnt_it(0);
    WeightOp wop{S};
    cub::TransformInputIterator<long long, WeightOp, cub::CountingInputIterator<int>> w_it(cnt_it, wop);
    size_t tb1 = 0, tb2 = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb1, w_it, wsum.p, n, s);
    cub::DoubleBuffer<unsigned long long> dk(keys.p, keys2.p);
    cub::DoubleBuffer<int> dv(rowsb.p, rows2.p);
    const int end_bit = kPartShift + bits_for((unsigned)nparts);
    cub::DeviceRadixSort::SortPairs(nullptr, tb2, dk, dv, n, 0, end_bit, s);
    DevBuf<unsigned char> tmp;
    tmp.alloc(std::max(tb1, tb2));
    size_t tb = tmp.n;
    if (pl.mode == 0) ILUB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, w_it, wsum.p, n, s));
    k_assign_parts<<<blocks_for(n, 256), 256, 0, s>>>(n, wsum.p, S, pl, part.p);
    k_sort_keys<<<blocks_for(n, 256), 256, 0, s>>>(n, S, part.p, T.level.p, keys.p, rowsb.p);
    tb = tmp.n;
    ILUB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, n, 0, end_bit, s));
    T.order.alloc(n);
    ILUB_CUDA(cudaMemcpyAsync(T.order.p, dv.Current(), sizeof(int) * n, cudaMemcpyDeviceToDevice, s));
    // chunk assignment (level-aligned when the (part, level) groups are big enough to fill chunks)
    const double avg_group = (double)n / ((double)std::max(1, T.nlevels) * (double)nparts);
    const int align = opt.level_align == 1 || (opt.level_align != 0 && avg_group >= (double)opt.align_min_rows) ? 1 : 0;
    T.level_aligned = align;
    DevBuf<int> gstart, newchunk, chunk_incl;
    gstart.alloc(n);
    newchunk.alloc(n);
    chunk_incl.alloc(n);
    k_group_start<<<blocks_for(n, 256), 256, 0, s>>>(n, dk.Current(), align, gstart.p);
    {
        size_t tbm = 0, tbs = 0;
        cub::DeviceScan::InclusiveScan(nullptr, tbm, gstart.p, gstart.p, cub::Max(), n, s);
        cub::DeviceScan::InclusiveSum(nullptr, tbs, newchunk.p, chunk_incl.p, n, s);
        DevBuf<unsigned char> tmpc;
        tmpc.alloc(std::max(tbm, tbs));
        size_t tbx = tmpc.n;
        ILUB_CUDA(cub::DeviceScan::InclusiveScan(tmpc.p, tbx, gstart.p, gstart.p, cub::Max(), n, s));
        k_new_chunk<<<blocks_for(n, 256), 256, 0, s>>>(n, gstart.p, newchunk.p);
        tbx = tmpc.n;
        ILUB_CUDA(cub::DeviceScan::InclusiveSum(tmpc.p, tbx, newchunk.p, chunk_incl.p, n, s));
        ILUB_CUDA(cudaStreamSynchronize(s));
    }
    T.part_chunk_ptr.alloc(nparts + 1);
    k_part_chunks<<<1, 256, 0, s>>>(n, nparts, dk.Current(), chunk_incl.p, T.part_chunk_ptr.p);
    int nchunks = 0;
    ILUB_CUDA(cudaMemcpyAsync(&nchunks, chunk_incl.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    T.nchunks = nchunks;
    const int64_t nlane = (int64_t)nchunks * 32;
    T.rows.alloc(nlane);
    T.rowpos.alloc(n);
    T.rlen.alloc(nlane);
    T.diag.alloc(nlane);
    T.diagpos.alloc(nlane);
    T.ckind.alloc(nchunks);
    T.chunk_ptr.alloc(nchunks + 1);
    ILUB_CUDA(cudaMemsetAsync(T.rows.p, 0xff, sizeof(int) * nlane, s));
    k_place_rows2<<<blocks_for(n, 256), 256, 0, s>>>(n, T.order.p, gstart.p, chunk_incl.p, T.rows.p, T.rowpos.p);
    // ---- chunk shapes
    DevBuf<int> cslabs;
    cslabs.alloc(nchunks + 1);
    ILUB_CUDA(cudaMemsetAsync(cslabs.p, 0, sizeof(int) * (nchunks + 1), s));
    ILUB_CUDA(cudaMemsetAsync(ctr.p, 0, sizeof(int) * 4, s));
    k_chunk_shape<<<blocks_for(nlane, 256), 256, 0, s>>>(nchunks, S, T.rows.p, (int)opt.wcap, T.rlen.p, cslabs.p,
                                                         T.ckind.p, ctr.p + 2, ctr.p + 3);
    size_t tb3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb3, cslabs.p, T.chunk_ptr.p, nchunks + 1, s);
    DevBuf<unsigned char> tmp3;
    tmp3.alloc(tb3);
    ILUB_CUDA(cub::DeviceScan::ExclusiveSum(tmp3.p, tb3, cslabs.p, T.chunk_ptr.p, nchunks + 1, s));
    int hshape[4];
    int nsl = 0;
    ILUB_CUDA(cudaMemcpyAsync(hshape, ctr.p, sizeof(hshape), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaMemcpyAsync(&nsl, T.chunk_ptr.p + nchunks, sizeof(int), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    T.maxw_a = hshape[2];
    T.any_kind_b = hshape[3];
    T.nslabs = nsl;
    {   // launch configuration of the sync-free kernel; `near` is baked into the encoded indices below
        int W = (int)std::min<int64_t>(16, std::max<int64_t>(2, opt.warps_per_cta));  // kernel: __launch_bounds__(512)
        int F = (int)std::min<int64_t>(W - 1, std::max<int64_t>(1, opt.fetch_warps));
        int RC = 32;
        while (RC * 2 <= opt.ring_chunks) RC *= 2;
        int near = (int)std::min<int64_t>(opt.near_window, RC / 2);
        int HR = 256;
        while (HR * 2 <= opt.halo_entries) HR *= 2;
        int stages = (int)std::min<int64_t>(2, std::max<int64_t>(1, opt.stages));
        const int slabs = std::max(1, T.maxw_a);
        const size_t budget = (size_t)227 * 1024 / (size_t)std::max<int64_t>(1, opt.ctas_per_sm) - 1024;
        auto need = [&]() {
            return (size_t)RC * 512 + 128 + (size_t)HR * 16 + (size_t)W * stages * slabs * 384 + (size_t)W * stages * 8 + (size_t)RC * 8 + 128 + 128;
        };
        while (need() > budget) {
            if (stages > 1) stages = 1;
            else if (RC > 64) { RC /= 2; near = std::min(near, RC / 2); }
            else if (HR > 256) HR /= 2;
            else if (W > 4) { W /= 2; F = std::min(F, W - 1); }
            else throw Error{ILUB_ERR_INVALID, "sync-free solve does not fit in shared memory (lower wcap)"};
        }
        T.cfg_warps = W;
        T.cfg_fetch = F;
        T.cfg_ring_chunks = RC;
        T.cfg_near = near;
        T.cfg_halo = HR;
        T.cfg_stages = stages;
        T.cfg_stage_slabs = slabs;
        T.cfg_smem = need();
    }
    const int64_t nent = (int64_t)nsl * 32;
    ILUB_REQUIRE(nent < (int64_t)1 << 31, ILUB_ERR_SIZE, "packed solve layout exceeds 2^31 entries");
    T.idx.alloc(nent);
    T.srcpos.alloc(nent);
    T.val.alloc(nent);
    T.gate.alloc(nchunks);
    DevBuf<unsigned char> pub;
    pub.alloc(n);
    ILUB_CUDA(cudaMemsetAsync(pub.p, 0, n, s));
    // pass 0: count remote edges per chunk -> halo pointers; pass 1: fill
    DevBuf<int> nrem;
    nrem.alloc(nchunks + 1);
    ILUB_CUDA(cudaMemsetAsync(nrem.p, 0, sizeof(int) * (nchunks + 1), s));
    k_chunk_fill<<<blocks_for(nlane, 256), 256, 0, s>>>(0, nchunks, S, T.rows.p, T.rlen.p, T.chunk_ptr.p, T.rowpos.p,
                                                        T.part_chunk_ptr.p, nparts, T.cfg_near, T.ckind.p, nullptr,
                                                        nullptr, nullptr, nullptr, T.level.p, nullptr, nrem.p, nullptr,
                                                        nullptr, nullptr, T.cfg_halo / 4);
    T.hptr.alloc(nchunks + 1);
    {
        size_t tbh = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tbh, nrem.p, T.hptr.p, nchunks + 1, s);
        DevBuf<unsigned char> tmph;
        tmph.alloc(tbh);
        ILUB_CUDA(cub::DeviceScan::ExclusiveSum(tmph.p, tbh, nrem.p, T.hptr.p, nchunks + 1, s));
        int nh = 0;
        ILUB_CUDA(cudaMemcpyAsync(&nh, T.hptr.p + nchunks, sizeof(int), cudaMemcpyDeviceToHost, s));
        ILUB_CUDA(cudaStreamSynchronize(s));
        T.nhalo = nh;
    }
    T.halo_src.alloc(T.nhalo);
    T.halo_cons.alloc(T.nhalo);
    k_chunk_fill<<<blocks_for(nlane, 256), 256, 0, s>>>(1, nchunks, S, T.rows.p, T.rlen.p, T.chunk_ptr.p, T.rowpos.p,
                                                        T.part_chunk_ptr.p, nparts, T.cfg_near, T.ckind.p, T.idx.p,
                                                        T.srcpos.p, T.diagpos.p, pub.p, T.level.p, T.gate.p, nrem.p,
                                                        T.hptr.p, T.halo_src.p, T.halo_cons.p, T.cfg_halo / 4);
    T.pubmask.alloc(nchunks);
    k_pubmask<<<blocks_for(nlane, 256), 256, 0, s>>>(nchunks, T.rows.p, pub.p, T.pubmask.p);
    T.mail.alloc(n);
    ILUB_CUDA(cudaMemsetAsync(T.mail.p, 0, sizeof(ulonglong2) * (size_t)n, s));
    T.epoch = 0;
    h->launches += 8;
    ILUB_CUDA(cudaStreamSynchronize(s));  // temporaries die here
    T.generic_built = true;
}

void pack_solve_values(ilub_handle* h, int slot, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    if (!T.built) return;
    if (T.sys.built) pack_systolic(h, slot, s);
    if (T.deep.built) pack_deep(h, slot, s);
    if (T.blk.built) pack_block(h, slot, s);
    if (T.lev.built) {
        if (T.lev.nent > 0)
            k_pack_values<<<blocks_for(T.lev.nent, 256), 256, 0, s>>>(T.lev.nent, T.lev.srcpos.p, h->lu.p, T.lev.val.p);
        k_pack_diag<<<blocks_for(h->n, 256), 256, 0, s>>>(h->n, T.lev.dpos.p, h->lu.p, T.lev.diag.p);
        h->launches += 2;
    }
    if (!T.generic_built) return;
    const int64_t nent = T.nslabs * 32;
    const int64_t nlane = (int64_t)T.nchunks * 32;
    if (nent > 0) k_pack_values<<<blocks_for(nent, 256), 256, 0, s>>>(nent, T.srcpos.p, h->lu.p, T.val.p);
    k_pack_diag<<<blocks_for(nlane, 256), 256, 0, s>>>(nlane, T.diagpos.p, h->lu.p, T.diag.p);
    h->launches += 2;
}

void build_level_sets(ilub_handle* h, int slot, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    if (T.lvl_built) return;
    const int n = h->n;
    T.lvl_ptr.alloc(T.nlevels + 1);
    T.lvl_rows.alloc(n);
    DevBuf<int> k1, k2, v1, v2;
    k1.alloc(n);
    k2.alloc(n);
    v1.alloc(n);
    v2.alloc(n);
    ILUB_CUDA(cudaMemcpyAsync(k1.p, T.level.p, sizeof(int) * n, cudaMemcpyDeviceToDevice, s));
    k_iota<<<blocks_for(n, 256), 256, 0, s>>>(n, v1.p);
    cub::DoubleBuffer<int> dk(k1.p, k2.p), dv(v1.p, v2.p);
    size_t tb = 0;
    const int eb = bits_for((unsigned)std::max(1, T.nlevels));
    cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, n, 0, eb, s);
    DevBuf<unsigned char> tmp;
    tmp.alloc(tb);
    ILUB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, n, 0, eb, s));
    ILUB_CUDA(cudaMemcpyAsync(T.lvl_rows.p, dv.Current(), sizeof(int) * n, cudaMemcpyDeviceToDevice, s));
    k_level_ptr<<<blocks_for(T.nlevels + 1, 256), 256, 0, s>>>(n, T.nlevels, dk.Current(), T.lvl_ptr.p);
    h->launches += 3;
    T.lvl_ptr_host.resize((size_t)T.nlevels + 1);
    ILUB_CUDA(cudaMemcpyAsync(T.lvl_ptr_host.data(), T.lvl_ptr.p, sizeof(int) * (T.nlevels + 1), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    T.lvl_built = true;
}

void build_level_packed(ilub_handle* h, int slot, cudaStream_t s) {
    TriSolve& T = h->solve[slot];
    if (T.lev.built) return;
    build_level_sets(h, slot, s);
    const int n = h->n;
    const SolveSource src = solve_source(h, slot);
    const SrcDev S = to_dev(src);
    DevBuf<int> cnt;
    cnt.alloc((size_t)n + 1);
    ILUB_CUDA(cudaMemsetAsync(cnt.p + n, 0, sizeof(int), s));
    k_lev_count<<<blocks_for(n, 256), 256, 0, s>>>(n, S, T.lvl_rows.p, cnt.p);
    T.lev.ptr.alloc((size_t)n + 1);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, T.lev.ptr.p, n + 1, s);
    DevBuf<unsigned char> tmp;
    tmp.alloc(tb);
    ILUB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, T.lev.ptr.p, n + 1, s));
    int nent = 0;
    ILUB_CUDA(cudaMemcpyAsync(&nent, T.lev.ptr.p + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    ILUB_CUDA(cudaStreamSynchronize(s));
    T.lev.nent = nent;
    T.lev.idx.alloc((size_t)std::max(nent, 1));
    T.lev.srcpos.alloc((size_t)std::max(nent, 1));
    T.lev.val.alloc((size_t)std::max(nent, 1));
    T.lev.dpos.alloc((size_t)n);
    T.lev.diag.alloc((size_t)n);
    k_lev_fill<<<blocks_for(n, 256), 256, 0, s>>>(n, S, src.unit_diag ? 1 : 0, T.lvl_rows.p, T.lev.ptr.p, T.lev.idx.p,
                                                  T.lev.srcpos.p, T.lev.dpos.p);
    h->launches += 3;
    ILUB_CUDA(cudaStreamSynchronize(s));  // temporaries die here
    T.lev.built = true;
}

void expand_batched(ilub_handle* h, const int* ptr1, const int* idx1, int32_t n1, int64_t nnz1, int32_t nsys,
                    cudaStream_t s) {
    const int64_t nnz = nnz1 * nsys;
    const int64_t n = (int64_t)n1 * nsys;
    const int64_t work = std::max(nnz, n + 1);
    k_expand_batched<<<blocks_for(work, 256), 256, 0, s>>>(n1, nnz1, nsys, ptr1, idx1, h->ptr.p, h->idx.p);
    h->launches++;
}

}  // namespace ilub
