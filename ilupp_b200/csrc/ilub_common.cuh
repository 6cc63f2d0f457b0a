// Internal declarations shared by the libilub translation units (not part of the C ABI).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "ilub.h"

namespace ilub {

// ------------------------------------------------------------------------------------------------
// errors: thrown inside the library, converted to status codes at the C ABI (ilub_capi.cu)
struct Error {
    int code;
    std::string msg;
};

#define ILUB_CUDA(expr)                                                                                  \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            throw ::ilub::Error{ILUB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)};     \
    } while (0)

#define ILUB_REQUIRE(cond, code, text)                      \
    do {                                                    \
        if (!(cond)) throw ::ilub::Error{(code), (text)};   \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device memory: a small caching allocator (ilub_capi.cu) in front of cudaMalloc / cudaFree.  The analysis phase
// allocates and frees dozens of temporaries; plain cudaMalloc / cudaFree (device-synchronising, milliseconds each on
// a busy allocator) made it 5-25x slower than its kernels.  A freed block is tagged with the stream the freeing
// thread works on (set_alloc_stream) and is only handed out again for work on the same stream, so stream order keeps
// a reused block safe without a device synchronisation.
void* pool_alloc(size_t bytes);
void pool_free(void* p) noexcept;
void pool_trim() noexcept;                 // return every cached block to the driver
void set_alloc_stream(cudaStream_t s);     // per thread: the stream subsequent allocations / frees are ordered on
// tag for blocks freed while the device is known to be idle (after a device synchronisation): reusable on any stream
static const cudaStream_t kIdleStream = (cudaStream_t)(intptr_t)-1;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) pool_free(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        p = static_cast<T*>(pool_alloc(count * sizeof(T)));
        n = count;
    }
    void ensure(size_t count) {
        if (count > n) alloc(count);
    }
    size_t bytes() const { return n * sizeof(T); }
};

// debug (env ILUB_DEBUG): wall-clock time of a host-side phase, stream synchronised at both ends
struct DebugTimer {
    const char* name;
    cudaStream_t s;
    std::chrono::steady_clock::time_point t0;
    bool on;
    DebugTimer(const char* name_, cudaStream_t s_) : name(name_), s(s_), on(getenv("ILUB_DEBUG") != nullptr) {
        if (on) {
            cudaStreamSynchronize(s);
            t0 = std::chrono::steady_clock::now();
        }
    }
    ~DebugTimer() {
        if (on) {
            cudaStreamSynchronize(s);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            fprintf(stderr, "[ilub] %-28s %8.3f ms\n", name, ms);
        }
    }
};

// per-thread, per-device "done once" flags for function attributes (cudaFuncSetAttribute is per device)
inline bool first_use_on_this_device(unsigned* masks /* one word per kernel variant */, int variant) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned bit = 1u << (dev & 31);
    if (masks[variant] & bit) return false;
    masks[variant] |= bit;
    return true;
}

struct Options {
    int64_t factor_algo = 0;    // 0 auto, 1 one launch per level, 2 sync-free, 3 level loop (one cooperative launch), 4 stencil (line-systolic; needs exact = 0)
    int64_t solve_algo = 0;     // 0 auto (line-systolic on structured grids, level loop on deep narrow DAGs, else sync-free), 1 level-scheduled, 2 sync-free, 3 level loop (one CTA), 4 cooperative level loop (grid barrier)
    int64_t exact = 1;          // IEEE mul/sub/div in reference order
    int64_t ctas_per_sm = 1;
    int64_t warps_per_cta = 16;
    int64_t near_window = 56;   // chunks; dependencies this close (same CTA) go through shared memory
    int64_t ring_chunks = 128;  // shared-memory x ring, in 32-row chunks (power of two, >= 2*near_window)
    int64_t wcap = 8;           // widest lane-per-row chunk (slabs); wider chunks are warp-cooperative
    int64_t fetch_warps = 2;    // warps per CTA that only move remote values from mailboxes into the halo ring
    int64_t halo_entries = 2048;// halo ring size (remote edges in flight per CTA)
    int64_t gate_slack = 8;     // start fine-grained polling this many chunks before the gate chunk completes
    int64_t stages = 2;         // TMA staging depth per warp
    int64_t spin_limit_ms = 1500;
    int64_t partition = 0;      // 0 auto (geometric strips/pencils when strides are detected), 1 contiguous ranges
    int64_t nparts = 0;         // 0 auto, else number of CTAs/parts to aim for
    int64_t part_kb = 256;      // auto nparts: aim for at least this many KB of solve data per CTA
    int64_t level_align = 2;    // chunks never mix levels: 0 never, 1 always, 2 when (part, level) groups average >= align_min_rows
    int64_t align_min_rows = 16;
    int64_t fastpath = 1;       // straight-line path for single-level lane-per-row chunks
    int64_t trace = 0;          // debug: record per-chunk clock stamps in the sync-free solve
    int64_t factor_group = 0;   // lanes per row in the factorisation kernels (0 = auto)
    int64_t sys_pz = 0;         // line-systolic: planes per CTA (0 = auto)
    int64_t sys_stages = 0;     // line-systolic: coefficient stages per warp (0 = auto)
    int64_t sys_pdl = 1;        // line-systolic: programmatic dependent launch (prologue of the next solve overlaps this one)
    int64_t sys_cluster = 0;    // line-systolic: largest thread-block cluster to try (0 = auto: 8, 1 = no clusters)
    int64_t sys_tile = 0;       // tile-systolic kernel (8 x 4 lines per warp) for plane-coupled 3-D grids: 0 off, 1 auto
    int64_t tile_fetch = 2;     // tile-systolic: mailbox fetcher warps per CTA
    int64_t tile_nap_fwd = 20, tile_nap_fetch = 20, tile_nap_tma = 20;  // tile-systolic: idle helper warps sleep this many ns between polls
    int64_t tile_nst = 0, tile_ring = 0;  // tile-systolic: coefficient stages / ring depth per warp (0 = auto)
    int64_t tile_debug = 0;     // tile-systolic: debug switches (timing experiments only)
    int64_t tile_warps = 0;     // tile-systolic: compute warps per CTA (0 = auto: 4, or 6 when the grid needs them)
    int64_t zero_copy = 0;      // apply() on a pinned host vector: the solve kernels read / write it directly over PCIe (1: both ways, 2: out only, 3: in only). Measured SLOWER than staged copies (64-byte PCIe requests): 0.91-1.06 ms against 0.79 ms at 128^3, so off
    int64_t deep_block = 1;     // deep narrow DAGs, fast arithmetic: block-sequential kernel (ilub_block.cu) instead of the level loop (0 = level loop)
    int64_t sys_plane = 1;      // plane-per-warp kernel for grids without a plane-to-plane dependency: 0 off, 1 auto, 2 whenever applicable
    int64_t plane_r = 1;        // plane-per-warp: right-hand sides per warp (1 or 2)
    int64_t levels_cluster = 1; // cooperative level loop: one thread-block cluster with the hardware barrier when the levels fit (0: always the grid barrier)
    int64_t plane_npc = 0;      // plane-per-warp: consecutive planes per CTA (0 = auto)
    int64_t plane_wpc = 0;      // plane-per-warp: consumer warps per CTA (0 = as many as shared memory holds, at most 8)
};
// The options in force for the calling thread: the process-wide defaults (ilub_set_option), or -- inside an entry point
// that works on a handle -- the defaults overlaid with that handle's own settings (ilub_set_handle_option, or options
// staged with ilub_stage_option before the handle was created).  Kernel-selection code only ever calls options().
Options& options();
struct OptionField {
    const char* name;
    int64_t Options::*member;
};
const OptionField* option_fields(int* count);
int option_index(const char* key);  // -1: unknown
void set_default_option(int field, int64_t value);
struct OptionScope {                // installs a handle's effective options for the lifetime of the scope
    explicit OptionScope(ilub_handle* h);
    ~OptionScope();
    Options* prev;
};

constexpr int kLanes = 32;
// dependency encoding inside a packed chunk (int32):
//   bit 31 set, bit 30 clear    local ring: low 30 bits = position of the producing row in this CTA's order
//   bit 31 set, bit 30 set      halo ring: low 30 bits = number of this remote edge in the CTA's edge order;
//                               the CTA's fetcher warps copy the producer's mailbox entry there
// Both are polled with one atomic 128-bit shared-memory load of {value, tag = position + 1}.
//   bit 31 clear                direct mailbox poll: original row index (chunks with too many remote edges
//                               for the halo ring; general path only)
constexpr unsigned kLocalBit = 0x80000000u;
constexpr unsigned kIntraBit = 0x40000000u;
constexpr unsigned kHaloBit = 0x40000000u;

// Line-systolic layout of one triangular solve (ilub_systolic.cu): chosen at analysis when every dependency of
// the solve sits at grid offset 1, nx or nx*ny of a lexicographically numbered nx x ny x nplanes grid.
constexpr int kSysStageSteps = 8;                            // steps per coefficient stage (8 KB per task)
constexpr long long kSysAbsentBits = 0x7ff8dead00000001ll;  // absent stencil entry (exact mode): a quiet NaN arithmetic never produces
struct SysSolve {
    bool built = false;
    bool tried = false;
    int nx = 0, ny = 0, nplanes = 0, nz1 = 1;  // nz1: planes per system (batches: nplanes = nz1 * nsys)
    int nyb = 0;                                // 32-line blocks per plane
    int nsteps_pad = 0;                         // (nx + 31) rounded up to the 8-step stage size
    int pyb = 1, pz = 1, gyb = 1, gz = 1;       // CTA tile of tasks (line blocks x planes) and CTA grid
    int csize = 1, gzp = 1;                     // thread-block cluster along z (DSMEM hand-offs), gz padded to it
    int zcoupled = 0;                           // planes depend on the previous plane (3-D)
    int order_xyz = 0;                          // exact arithmetic: accumulate x, y, z (bwd-gather) instead of z, y, x
    int ntasks = 0, ntasks1 = 0;                // tasks = (line block, plane); ntasks1: tasks of one system
    int nstages = 3, rz = 8, ry = 64;
    // tile-systolic variant (ilub_tile.cu): a warp owns ty x tz adjacent lines (lane = a + ty * b, skew a + b), so the
    // y AND the z dependency of most rows is a shuffle.  The line-systolic layout above is the case ty = 32, tz = 1.
    int tile = 0, ty = 32, tz = 1;
    int nty = 0, ntz = 0, ntz1 = 0;             // tiles along y, along z (all systems), along z per system
    int twy = 2, twz = 2, tgy = 1, tgz = 1;     // tiles per CTA and CTA grid
    int tccy = 0, tccz = 0;                     // thread-block cluster shape in CTAs (0: not chosen yet)
    int tnst = 3, tnwin = 6, tring = 64;        // coefficient stages, window blocks, hand-off ring depth (steps) per warp
    size_t smem = 0;
    int nzchan = 0, nychan = 0;
    int packed_exact = -1;                      // arithmetic mode the coefficients were packed for
    DevBuf<int> srcpos;                         // [ntasks1][nsteps_pad][4][32] position in lu (-1: absent / unit)
    DevBuf<double> coef;                        // [ntasks][nsteps_pad][4][32] {z, y, x, diagonal} per lane and step
    DevBuf<ulonglong2> zmail, ymail;            // cross-CTA hand-off {x, epoch}
    // stencil factorisation on the same kernel (slot 0 only): value / partner positions and the product stream
    bool ftried = false, fbuilt = false;
    DevBuf<int> fsrc, fpart;
    DevBuf<int> frow;        // [nnz of one system] row of every entry of the CSR-read arrays (k_stencil_finish works per entry)
    DevBuf<double> fcoef;
    unsigned long long epoch = 0;
};

// Level-loop layout of one triangular solve for deep, narrow DAGs (ilub_deep.cu): everything in solve order.
struct DeepSolve {
    bool built = false;
    int nent = 0;
    int packed_exact = -1;
    DevBuf<int> eptr;       // [n + 1] first packed entry of every solve position
    DevBuf<int4> meta;      // [n] {row, first entry, length, position of the diagonal in lu (-1: unit)}
    DevBuf<int> dpos;       // [nent] solve position of the dependency
    DevBuf<int> dsrc;       // [nent] position of the coefficient in lu
    DevBuf<double> val;     // [nent] coefficient (pre-divided by the diagonal unless exact)
    DevBuf<double> diag;    // [n] diagonal (exact) or its reciprocal
    DevBuf<double> bs, xs;  // [n] right-hand side / solution in solve order
};

// Block-sequential layout (ilub_block.cu: deep, narrow, band-local DAGs in the fast arithmetic mode): rows in their
// natural solve order in blocks of 32; per row a "far" list (dependencies older than 512 rows, reduced ahead of time by
// stager warps) and a "near" list (recent entries, then the entries inside the row's own block)
struct BlockSolve {
    bool built = false, packed = false;
    int nblocks = 0, nfar = 0, nnear = 0;
    DevBuf<int> fptr, fpos, fsrc;   // far lists: [n + 1], dependency position, coefficient source in lu
    DevBuf<int> nptr, npos, nsrc;   // near lists
    DevBuf<int> cr;                 // [n] recent entries per row (the rest of its near list is local)
    DevBuf<int> co;                 // [n] ... of which older than the previous block (first in the list)
    DevBuf<int> dsrc;               // [n] diagonal source in lu (-1: unit)
    DevBuf<double> fval, nval, rdiag;
    DevBuf<int4> bmeta;             // [nblocks] {depth | leaf << 16 | widest local list << 20, widest recent lists, long-local mask, long-recent mask}
    DevBuf<int> llev;               // [n] level of every row inside its block's local DAG (the block factorisation walks them)
    DevBuf<int> bdepth;             // [nblocks] deepest local level of the block (all rows)
    // the block's local system solved ahead of time: row i of (I + L_local)^-1 - I on the transitive closure of the row's
    // local dependencies, so that the solve needs ONE exchange of values inside the block instead of depth(block) sweeps
    int nclo = 0;
    DevBuf<int> cptr, cpos;         // [n + 1], position inside the block (0 .. 31)
    DevBuf<double> cval;
};

// One triangular solve packed in solve order (see DESIGN.md "data layout in HBM").
struct TriSolve {
    bool built = false;
    int n = 0;
    int nparts = 0;
    int nchunks = 0;
    int64_t nslabs = 0;     // 32-entry slabs of idx/val
    int maxw_a = 0;         // widest lane-per-row chunk
    int nlevels = 0;
    bool backward = false;
    bool unit_diag = false;
    int any_kind_b = 0;
    DevBuf<int> level;            // [n] level of every row in this solve's DAG
    DevBuf<int> order;            // [n] rows sorted by (part, level, row)
    DevBuf<int> rowpos;           // [n] (chunk << 5 | lane) of every row
    DevBuf<int> part_chunk_ptr;   // [nparts + 1]
    DevBuf<int> chunk_ptr;        // [nchunks + 1] slab offset of every chunk
    DevBuf<unsigned char> ckind;  // [nchunks] bit0: warp-cooperative (kind B); bit1: has same-chunk deps; bit2: has mailbox deps; bit3: single level; bit4: direct mailbox polls
    DevBuf<int> rows;             // [nchunks * 32] original row per lane (-1 = padding)
    DevBuf<int> rlen;             // [nchunks * 32] number of off-diagonal entries per lane
    DevBuf<int> idx;              // [nslabs * 32] encoded dependencies
    DevBuf<int> srcpos;           // [nslabs * 32] position of the value in the factor value array (-1 pad)
    DevBuf<double> val;           // [nslabs * 32]
    DevBuf<double> diag;          // [nchunks * 32]
    DevBuf<int> diagpos;          // [nchunks * 32] position of the diagonal value (-1: unit / padding)
    DevBuf<int> gate;             // [nchunks] gate row: last local position feeding this chunk's lowest level (-1: none)
    DevBuf<unsigned> pubmask;     // [nchunks] lanes whose row is read through the global path by someone
    DevBuf<int> hptr;             // [nchunks + 1] first remote edge (halo position) of every chunk
    DevBuf<int> halo_src;         // [nhalo] producing row of every remote edge
    DevBuf<int> halo_cons;        // [nhalo] consuming chunk (CTA numbering) of every remote edge
    int nhalo = 0;
    DevBuf<ulonglong2> mail;      // [n] {x bits, epoch tag}: the global dependency path (one 16-B atomic access)
    unsigned long long epoch = 0;
    int level_aligned = 0;
    int part_mode = 0;            // 0 contiguous ranges, 2 / 3: strips / pencils from detected strides, 4 line-systolic, 5 level loop (one CTA), 6 cooperative level loop, 7 one launch per level (few wide levels)
    int stride_y = 0, stride_z = 0, grid_y = 1, grid_z = 1;
    // launch configuration of the sync-free kernel, fixed at analysis (near is baked into idx)
    int cfg_warps = 16, cfg_fetch = 2, cfg_ring_chunks = 128, cfg_near = 56, cfg_halo = 2048, cfg_stages = 2,
        cfg_stage_slabs = 1;
    size_t cfg_smem = 0;
    // level-set listing for the level-scheduled variant / ilub_get_level_sets (built on demand)
    DevBuf<long long> trace;      // debug stamps (option trace)
    DevBuf<int> lvl_ptr, lvl_rows;
    std::vector<int> lvl_ptr_host;  // host copy of lvl_ptr (the one-launch-per-level solve sizes its grids from it)
    bool lvl_built = false;
    bool generic_built = false;   // the chunked sync-free layout above exists (always for slot 0: the factorisation uses it)
    // level-packed layout (few very wide levels, part_mode 7): rows in (level, row) order = lvl_rows; their entries,
    // in the order the reference consumes them, packed contiguously in the same order
    struct LevPacked {
        bool built = false;
        int64_t nent = 0;
        DevBuf<int> ptr;     // [n + 1] first entry of every position
        DevBuf<int> idx;     // [nent] dependency (original row index)
        DevBuf<int> srcpos;  // [nent] position of the coefficient in the factor value array
        DevBuf<int> dpos;    // [n] position of the diagonal (-1: unit diagonal)
        DevBuf<double> val;  // [nent]
        DevBuf<double> diag; // [n]
    } lev;
    SysSolve sys;                 // line-systolic layout (structured grids), preferred when built
    DeepSolve deep;               // level-loop layout (deep narrow DAGs)
    BlockSolve blk;               // block-sequential layout (deep narrow band-local DAGs, fast arithmetic)
};

// Transposed pattern of the strict lower or strict upper triangle (for scatter-form solves).
struct Transposed {
    bool built = false;
    DevBuf<int> tptr;  // [n + 1]
    DevBuf<int> tidx;  // [nt] original row k of entry (k, i), ascending k inside every i
    DevBuf<int> tpos;  // [nt] position of that entry in the CSR value array
    int64_t nt = 0;
};

// How to read one triangular solve out of the CSR arrays (ptr, idx, lu, dpos).
struct SolveSource {
    const int* beg_a = nullptr;  // entries of row i: [beg_a[i] + beg_off, end_a[i] + end_off)
    int beg_off = 0;
    const int* end_a = nullptr;
    int end_off = 0;
    const int* dep_idx = nullptr;  // dependency row per entry
    const int* val_pos = nullptr;  // position of the value in lu per entry (nullptr: the entry position itself)
    const int* diag_pos = nullptr; // dpos (nullptr with unit_diag)
    bool unit_diag = false;
    bool backward = false;         // rows resolve in descending order
    bool reverse_entries = false;  // accumulate the stored entries last-to-first (bwd-scatter order)
};

// device-side view of a SolveSource
struct SrcDev {
    const int* beg_a;
    int beg_off;
    const int* end_a;
    int end_off;
    const int* dep_idx;
    const int* val_pos;
    const int* diag_pos;
    int reverse;
    int backward;
};
#ifdef __CUDACC__
__host__ __device__ __forceinline__ int src_beg(const SrcDev& S, int i) { return S.beg_a[i] + S.beg_off; }
__host__ __device__ __forceinline__ int src_end(const SrcDev& S, int i) { return S.end_a[i] + S.end_off; }
#endif
inline SrcDev to_dev(const SolveSource& s) {
    SrcDev d;
    d.beg_a = s.beg_a;
    d.beg_off = s.beg_off;
    d.end_a = s.end_a;
    d.end_off = s.end_off;
    d.dep_idx = s.dep_idx;
    d.val_pos = s.val_pos;
    d.diag_pos = s.diag_pos;
    d.reverse = s.reverse_entries ? 1 : 0;
    d.backward = s.backward ? 1 : 0;
    return d;
}

enum Kind { KIND_ILU0 = 0, KIND_ICHOL0 = 1 };

}  // namespace ilub

// The opaque handle of the C ABI.
struct ilub_handle {
    int kind = 0;        // ilub::Kind
    int is_csr = 1;      // orientation label of the input (CSC input = CSR arrays of A^T)
    int device = 0;
    int sm_count = 148;
    long long clock_khz = 1965000;  // SM clock (cycles per ms) for the watchdog limits
    int32_t n = 0;       // dimension (nsys * n_one for batches)
    int64_t nnz = 0;
    int32_t nsys = 1;
    int prefactored = 0;  // created from factors computed elsewhere (ilub_create_from_factors): `a` holds the factors, not a matrix
    // the matrix as given (CSR-read arrays), kept for spmv / bicgstab / refactor
    ilub::DevBuf<int> ptr, idx;
    ilub::DevBuf<double> a;
    // factor values on A's pattern (ILU0: L strictly below, U on/above; IChol0: L on tril, rest unused)
    ilub::DevBuf<double> lu;
    ilub::DevBuf<int> dpos;  // [n] position of the diagonal entry of every row
    int64_t nnz_lower = 0;   // strictly lower entries
    int64_t nnz_upper = 0;   // strictly upper entries
    ilub::Transposed t_upper, t_lower;
    // solves: [0] first of apply, [1] second of apply, [2] first of apply_trans, [3] second of apply_trans
    ilub::TriSolve solve[4];
    ilub::DevBuf<unsigned char> fflag;  // factorisation row flags
    ilub::DevBuf<unsigned> fbarrier;    // grid barrier of the level-loop factorisation
    ilub::DevBuf<unsigned> fblock;      // hand-shake words of the block-sequential factorisation (k_factor_block)
    unsigned fblock_epoch = 0;
    ilub::DevBuf<int> fnfar;            // [n] far eliminations per row (k_factor_block)
    ilub::DevBuf<double> fpivot;        // pivots of the stencil factorisation
    int max_row_len = -1;               // longest row of the CSR-read arrays (computed once, factor_prepare)
    unsigned fepoch = 0;
    ilub::DevBuf<int> status;           // device status word(s) written by kernels (watchdog / structure errors)
    ilub::DevBuf<double> work;          // vector workspace (host-side apply, bicgstab)
    // pivoting permutations of factors computed elsewhere (ilub_set_permutations): apply = scatter_out . T^-1 . gather_in
    ilub::DevBuf<int> perm_in, perm_out;
    ilub::DevBuf<double> permbuf;
    ilub::DevBuf<double> xalign;        // 16-byte aligned stand-in for an odd-aligned vector (tile-systolic tensor maps)
    ilub::DevBuf<double> scal;          // device scalars for bicgstab
    // line-systolic CTA tiling of the last plane-coupled solve analysed (the solves of one matrix share the grid, and
    // the occupancy queries behind the choice cost more than the rest of the analysis): key -> {pz, nst, rz, csize, gzp}
    long long sys_tiling_key[4] = {-1, -1, -1, -1};
    int sys_tiling[5] = {0, 0, 0, 0, 0};
    // per-handle option overrides {field, value} and the effective set while an entry point works on the handle
    std::vector<std::pair<int, int64_t>> opt_overrides;
    ilub::Options opt_effective;
    double timings[4] = {0, 0, 0, 0};
    int64_t launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // The handle is not re-entrant (shared workspace, mailbox epochs, lazily built solves): every entry point of the
    // C ABI that touches it holds this lock, so threads sharing one preconditioner are serialised like the
    // reference's GIL-holding apply (binding.cpp:237-245) serialises them.
    std::recursive_mutex mu;
    // Device-path calls (where = ILUB_DEVICE) are asynchronous: a watchdog abort of one of their kernels is copied
    // here behind the kernels (pinned host word) and raised by the NEXT call on the handle / ilub_check_status.
    int* host_status = nullptr;
    ~ilub_handle() {
        for (int i = 0; i < 4; ++i)
            if (ev[i]) cudaEventDestroy(ev[i]);
        if (host_status) cudaFreeHost(host_status);
    }
};

namespace ilub {

// ---- analysis (ilub_analysis.cu)
void validate_and_find_diag(ilub_handle* h, cudaStream_t s);
void build_transposed(ilub_handle* h, bool upper, cudaStream_t s);
SolveSource solve_source(ilub_handle* h, int which);
void analyse_solve(ilub_handle* h, int which, cudaStream_t s);            // levels + layout(s)
void analyse_generic(ilub_handle* h, int which, cudaStream_t s);          // partition + chunked sync-free layout
void pack_solve_values(ilub_handle* h, int which, cudaStream_t s);        // (re)fill val/diag from lu
void build_level_sets(ilub_handle* h, int which, cudaStream_t s);
void build_level_packed(ilub_handle* h, int slot, cudaStream_t s);  // level-packed layout of a few-wide-levels solve
bool deep_wanted(const ilub_handle* h, int slot);
bool block_wanted(const ilub_handle* h, int slot);   // block-sequential solve (ilub_block.cu)
void analyse_block(ilub_handle* h, int slot, cudaStream_t s);
void pack_block(ilub_handle* h, int slot, cudaStream_t s);
void run_block(ilub_handle* h, int slot, double* x, int64_t ldx, int nrhs, cudaStream_t s);
int levels_block(ilub_handle* h, int slot, cudaStream_t s);  // dependency levels through the block layout (ilub_block.cu)
bool factor_block(ilub_handle* h, cudaStream_t s);  // block-sequential ILU(0) for deep narrow DAGs (ilub_factor.cu); false: not applicable
bool levels_wanted(const ilub_handle* h, int slot);   // cooperative level loop (ilub_trsv.cu)
bool few_levels_wanted(const ilub_handle* h, int slot);  // one launch per level (ilub_trsv.cu)
void analyse_deep(ilub_handle* h, int slot, cudaStream_t s);
void pack_deep(ilub_handle* h, int slot, cudaStream_t s);
void run_deep(ilub_handle* h, int slot, double* x, int64_t ldx, int nrhs, cudaStream_t s);
bool stencil_factor_possible(ilub_handle* h, cudaStream_t s);
void factor_stencil(ilub_handle* h, cudaStream_t s);
int factor_schedule(const ilub_handle* h);
void factor_prepare(ilub_handle* h, cudaStream_t s);
int read_status(ilub_handle* h, cudaStream_t s);                          // syncs the stream
void expand_batched(ilub_handle* h, const int* ptr1, const int* idx1, int32_t n1, int64_t nnz1, int32_t nsys,
                    cudaStream_t s);

// ---- numeric factorisation (ilub_factor.cu)
void factor_numeric(ilub_handle* h, cudaStream_t s);

// ---- line-systolic triangular solves on structured grids (ilub_systolic.cu)
bool analyse_systolic(ilub_handle* h, int which, cudaStream_t s);         // false: structure not eligible
void pack_systolic(ilub_handle* h, int which, cudaStream_t s);
void run_systolic(ilub_handle* h, int which, double* x, int64_t ldx, int nrhs, cudaStream_t s);
bool systolic_out_of_place_ok(const ilub_handle* h, int which);
void run_systolic_io(ilub_handle* h, int which, const double* xin, double* xout, cudaStream_t s);
// ---- tile-systolic variant for plane-coupled 3-D grids (ilub_tile.cu)
bool tile_configure(ilub_handle* h, SysSolve& Y, int nsys);               // false: this grid does not fit the tile kernel
void launch_tile(ilub_handle* h, int which, double* x, int64_t ldx, int nrhs, cudaStream_t s, const double* fcoef);
// ---- plane-per-warp variant for 2-D grids / batches and many right-hand sides (ilub_plane.cu); false: not applicable
bool run_plane(ilub_handle* h, int which, double* x, int64_t ldx, int nrhs, cudaStream_t s);

// ---- triangular solves (ilub_trsv.cu)
void run_solve(ilub_handle* h, int which, double* x, int64_t ldx, int nrhs, cudaStream_t s);
void extract_factor(ilub_handle* h, int which, double* data, int* indices, int* indptr, cudaStream_t s);

// ---- permutations around the solves of pivoted factors (ilub_krylov.cu)
void permute_gather(ilub_handle* h, const int* perm, double* x, int64_t ldx, int nrhs, cudaStream_t s);   // x[k] <- x[perm[k]]
void permute_scatter(ilub_handle* h, const int* perm, double* x, int64_t ldx, int nrhs, cudaStream_t s);  // x[perm[k]] <- x[k]

// ---- Krylov (ilub_krylov.cu)
void spmv(ilub_handle* h, const double* x, double* y, cudaStream_t s);
int bicgstab(ilub_handle* h, const double* b, double* x, double rtol, double atol, int max_iter, int* iters,
             double* rel, double* abs_, cudaStream_t s);
int pcg(ilub_handle* h, const double* b, double* x, double rtol, double atol, int max_iter, int* iters, double* rel,
        double* abs_, cudaStream_t s);

}  // namespace ilub
