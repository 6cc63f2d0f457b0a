// The C ABI of libilub.so (include/ilub.h): argument checking, host<->device staging, error
// translation.  No numerical work happens here.
#include <chrono>
#include <cstring>
#include <memory>
#include <vector>

#include <cstdio>

#include "ilub_common.cuh"

using namespace ilub;

namespace {

thread_local std::string g_last_error;

template <class F>
int guarded(F&& f) {
    try {
        f();
        return ILUB_OK;
    } catch (const Error& e) {
        g_last_error = e.msg;
        return e.code;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return ILUB_ERR_INVALID;
    } catch (...) {
        g_last_error = "unknown error";
        return ILUB_ERR_INVALID;
    }
}

cudaMemcpyKind to_dev(int where) { return where == ILUB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice; }
cudaMemcpyKind from_dev(int where) { return where == ILUB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost; }

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        ILUB_CUDA(cudaGetDevice(&prev));
        if (prev != dev) ILUB_CUDA(cudaSetDevice(dev));
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// slots of the two solves of apply (transpose = 0) / apply_trans (transpose = 1), in order
void apply_slots(const ilub_handle* h, int transpose, int& first, int& second) {
    if (h->kind == KIND_ICHOL0) {  // LL^T class: L then L^T either way (preconditioner_implementation.h:381-394)
        first = 0;
        second = 3;
        return;
    }
    const bool gather = (h->is_csr != 0) != (transpose != 0);
    if (gather) {  // fwd-gather(Lc), bwd-gather(Uc)
        first = 0;
        second = 1;
    } else {       // fwd-scatter(Uc) = Uc^T forward, bwd-scatter(Lc) = Lc^T backward
        first = 2;
        second = 3;
    }
}

void ensure_solve(ilub_handle* h, int slot, cudaStream_t s) {
    if (h->solve[slot].built) return;
    analyse_solve(h, slot, s);
    pack_solve_values(h, slot, s);
}

typedef std::lock_guard<std::recursive_mutex> HandleLock;

// options staged by this thread for the handle it creates next (ilub_stage_option)
thread_local std::vector<std::pair<int, int64_t>> tl_staged_options;

// a kernel of an earlier asynchronous (device-path) call gave up: raise it now instead of returning garbage silently
void raise_deferred(ilub_handle* h, cudaStream_t s) {
    if (!h->host_status || *(volatile int*)h->host_status == 0) return;
    const int code = *(volatile int*)h->host_status;
    *(volatile int*)h->host_status = 0;
    ILUB_CUDA(cudaMemsetAsync(h->status.p, 0, sizeof(int) * 4, s));
    throw Error{ILUB_ERR_WATCHDOG, "an earlier asynchronous call on this handle was aborted by the watchdog (code " +
                                       std::to_string(code) + "): its results are invalid"};
}
void post_deferred(ilub_handle* h, cudaStream_t s) {
    if (h->host_status) ILUB_CUDA(cudaMemcpyAsync(h->host_status, h->status.p, sizeof(int), cudaMemcpyDeviceToHost, s));
}

// device-side address of a HOST pointer if it is pinned (page-locked) memory the current device can access, else null
double* device_view_of_host(const double* x) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, x) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return nullptr;
    return static_cast<double*>(at.devicePointer);
}

int slot_of(const ilub_handle* h, int which_solve) {
    int a, b;
    apply_slots(h, which_solve >= 2 ? 1 : 0, a, b);
    return (which_solve & 1) ? b : a;
}

void create_common(int kind, const double* data, const int32_t* indices, const int32_t* indptr, int32_t n1,
                   int64_t nnz1, int32_t nsys, int is_csr, int where, int device, void* stream, ilub_handle** out,
                   bool prefactored = false) {
    // options staged by this thread belong to this creation, whether it succeeds or not
    std::vector<std::pair<int, int64_t>> staged;
    staged.swap(tl_staged_options);
    ILUB_REQUIRE(out != nullptr, ILUB_ERR_INVALID, "out is null");
    *out = nullptr;
    ILUB_REQUIRE(data && indices && indptr, ILUB_ERR_INVALID, "null matrix array");
    ILUB_REQUIRE(n1 > 0, ILUB_ERR_INVALID, "matrix has size 0!");  // binding.cpp:80-81
    ILUB_REQUIRE(nnz1 >= 0 && nsys >= 1, ILUB_ERR_INVALID, "bad nnz / nsys");
    ILUB_REQUIRE(where == ILUB_HOST || where == ILUB_DEVICE, ILUB_ERR_INVALID, "bad `where`");
    const int64_t n64 = (int64_t)n1 * nsys, nnz64 = nnz1 * nsys;
    ILUB_REQUIRE(n64 < ((int64_t)1 << 31) - 64 && nnz64 < ((int64_t)1 << 31) - 64, ILUB_ERR_SIZE,
                 "problem exceeds 32-bit indexing (reference Integer is int32, declarations.h:49-53)");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    ILUB_REQUIRE(ce == cudaSuccess && ndev > 0, ILUB_ERR_CUDA,
                 std::string("no CUDA device available (libilub has no CPU fallback): ") + cudaGetErrorString(ce));
    ILUB_REQUIRE(device >= 0 && device < ndev, ILUB_ERR_INVALID, "bad device ordinal");
    DeviceGuard guard(device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    set_alloc_stream(s);
    std::unique_ptr<ilub_handle> h(new ilub_handle());
    h->opt_overrides.swap(staged);
    OptionScope scope(h.get());
    h->kind = kind;
    h->is_csr = is_csr ? 1 : 0;
    h->device = device;
    h->n = (int32_t)n64;
    h->nnz = nnz64;
    h->nsys = nsys;
    h->prefactored = prefactored ? 1 : 0;
    cudaDeviceProp prop;
    ILUB_CUDA(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    h->clock_khz = prop.clockRate > 0 ? prop.clockRate : 1965000;
    for (int i = 0; i < 4; ++i) ILUB_CUDA(cudaEventCreate(&h->ev[i]));
    ILUB_CUDA(cudaHostAlloc((void**)&h->host_status, sizeof(int), cudaHostAllocDefault));
    *h->host_status = 0;
    h->ptr.alloc(n64 + 1);
    h->idx.alloc(nnz64);
    h->a.alloc(nnz64);
    h->lu.alloc(nnz64);
    if (nsys == 1) {
        ILUB_CUDA(cudaMemcpyAsync(h->ptr.p, indptr, sizeof(int) * (n64 + 1), to_dev(where), s));
        ILUB_CUDA(cudaMemcpyAsync(h->idx.p, indices, sizeof(int) * nnz64, to_dev(where), s));
    } else {
        DevBuf<int> p1, i1;
        p1.alloc(n1 + 1);
        i1.alloc(nnz1);
        ILUB_CUDA(cudaMemcpyAsync(p1.p, indptr, sizeof(int) * (n1 + 1), to_dev(where), s));
        ILUB_CUDA(cudaMemcpyAsync(i1.p, indices, sizeof(int) * nnz1, to_dev(where), s));
        // the single-system pointer array must be consistent with nnz1 before we replicate it
        int last = 0;
        ILUB_CUDA(cudaMemcpyAsync(&last, p1.p + n1, sizeof(int), cudaMemcpyDeviceToHost, s));
        ILUB_CUDA(cudaStreamSynchronize(s));
        ILUB_REQUIRE(last == nnz1, ILUB_ERR_INVALID, "indptr[n] != nnz");
        expand_batched(h.get(), p1.p, i1.p, n1, nnz1, nsys, s);
        ILUB_CUDA(cudaStreamSynchronize(s));
    }
    ILUB_CUDA(cudaMemcpyAsync(h->a.p, data, sizeof(double) * nnz64, to_dev(where), s));
    if (nsys == 1) {
        int ends[1] = {0};
        ILUB_CUDA(cudaMemcpyAsync(ends, h->ptr.p + n64, sizeof(int), cudaMemcpyDeviceToHost, s));
        ILUB_CUDA(cudaStreamSynchronize(s));
        ILUB_REQUIRE(ends[0] == nnz64, ILUB_ERR_INVALID, "indptr[n] != nnz");
    }

    ILUB_CUDA(cudaEventRecord(h->ev[0], s));
    {
        DebugTimer t("validate + diagonal", s);
        validate_and_find_diag(h.get(), s);
    }
    int first, second;
    apply_slots(h.get(), 0, first, second);
    {
        DebugTimer t("analyse slot 0", s);
        analyse_solve(h.get(), 0, s);  // the factorisation DAG (= L solve DAG) is always needed
    }
    {
        DebugTimer t("analyse first solve", s);
        analyse_solve(h.get(), first, s);
    }
    {
        DebugTimer t("analyse second solve", s);
        analyse_solve(h.get(), second, s);
    }
    if (!prefactored) {
        DebugTimer t("factorisation schedule", s);
        factor_prepare(h.get(), s);
    }
    ILUB_CUDA(cudaEventRecord(h->ev[1], s));
    if (prefactored) {
        // the values ARE the factors (computed elsewhere): nothing to factorise
        ILUB_CUDA(cudaMemcpyAsync(h->lu.p, h->a.p, sizeof(double) * nnz64, cudaMemcpyDeviceToDevice, s));
    } else {
        DebugTimer t("numeric factorisation", s);
        factor_numeric(h.get(), s);
    }
    ILUB_CUDA(cudaEventRecord(h->ev[2], s));
    DebugTimer tpack("pack", s);
    for (int k = 0; k < 4; ++k) pack_solve_values(h.get(), k, s);
    ILUB_CUDA(cudaEventRecord(h->ev[3], s));
    read_status(h.get(), s);  // syncs; throws on watchdog
    float ms01 = 0, ms12 = 0, ms23 = 0;
    cudaEventElapsedTime(&ms01, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&ms12, h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&ms23, h->ev[2], h->ev[3]);
    h->timings[0] = ms01;
    h->timings[1] = ms12;
    h->timings[2] = ms23;
    h->timings[3] = ms01 + ms12 + ms23;
    *out = h.release();
}

}  // namespace

extern "C" {

const char* ilub_last_error(void) { return g_last_error.c_str(); }
int ilub_version(void) { return 100; }
int ilub_index_size(void) { return (int)sizeof(int32_t); }

int ilub_device_count(int* count) {
    return guarded([&] {
        ILUB_REQUIRE(count, ILUB_ERR_INVALID, "count is null");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) {
            *count = 0;
            throw Error{ILUB_ERR_CUDA, cudaGetErrorString(e)};
        }
        *count = n;
    });
}

int ilub_create_ilu0(const double* data, const int32_t* indices, const int32_t* indptr, int32_t n, int64_t nnz,
                     int is_csr, int where, int device, void* stream, ilub_handle** out) {
    return guarded([&] { create_common(KIND_ILU0, data, indices, indptr, n, nnz, 1, is_csr, where, device, stream, out); });
}
int ilub_create_ichol0(const double* data, const int32_t* indices, const int32_t* indptr, int32_t n, int64_t nnz,
                       int is_csr, int where, int device, void* stream, ilub_handle** out) {
    return guarded(
        [&] { create_common(KIND_ICHOL0, data, indices, indptr, n, nnz, 1, is_csr, where, device, stream, out); });
}
int ilub_create_ilu0_batched(const double* values, const int32_t* indices, const int32_t* indptr, int32_t n,
                             int64_t nnz, int32_t nsys, int is_csr, int where, int device, void* stream,
                             ilub_handle** out) {
    return guarded(
        [&] { create_common(KIND_ILU0, values, indices, indptr, n, nnz, nsys, is_csr, where, device, stream, out); });
}
int ilub_create_ichol0_batched(const double* values, const int32_t* indices, const int32_t* indptr, int32_t n,
                               int64_t nnz, int32_t nsys, int is_csr, int where, int device, void* stream,
                               ilub_handle** out) {
    return guarded([&] {
        create_common(KIND_ICHOL0, values, indices, indptr, n, nnz, nsys, is_csr, where, device, stream, out);
    });
}

int ilub_create_from_factors(int kind, const double* data, const int32_t* indices, const int32_t* indptr, int32_t n,
                             int64_t nnz, int is_csr, int where, int device, void* stream, ilub_handle** out) {
    return guarded([&] {
        ILUB_REQUIRE(kind == KIND_ILU0 || kind == KIND_ICHOL0, ILUB_ERR_INVALID, "kind must be 0 (L U) or 1 (L L^T)");
        create_common(kind, data, indices, indptr, n, nnz, 1, is_csr, where, device, stream, out, true);
    });
}

int ilub_set_permutations(ilub_handle* h, const int32_t* gather_in, const int32_t* scatter_out, int where, void* stream) {
    return guarded([&] {
        ILUB_REQUIRE(h, ILUB_ERR_INVALID, "null handle");
        ILUB_REQUIRE(h->prefactored, ILUB_ERR_INVALID, "permutations belong to factors computed elsewhere (ilub_create_from_factors)");
        HandleLock lock(h->mu);
        OptionScope scope(h);
        DeviceGuard guard(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        set_alloc_stream(s);
        auto load = [&](DevBuf<int>& buf, const int32_t* p) {
            buf.release();
            if (!p) return;
            if (where == ILUB_HOST) {  // a permutation: every index in range, none twice
                std::vector<char> seen((size_t)h->n, 0);
                for (int32_t k = 0; k < h->n; ++k) {
                    ILUB_REQUIRE(p[k] >= 0 && p[k] < h->n && !seen[p[k]], ILUB_ERR_INVALID, "not a permutation");
                    seen[p[k]] = 1;
                }
            }
            buf.alloc((size_t)h->n);
            ILUB_CUDA(cudaMemcpyAsync(buf.p, p, sizeof(int) * (size_t)h->n, to_dev(where), s));
        };
        load(h->perm_in, gather_in);
        load(h->perm_out, scatter_out);
        ILUB_CUDA(cudaStreamSynchronize(s));
    });
}

int ilub_refactor(ilub_handle* h, const double* data, int where, void* stream) {
    return guarded([&] {
        ILUB_REQUIRE(h && data, ILUB_ERR_INVALID, "null argument");
        HandleLock lock(h->mu);
        OptionScope scope(h);
        DeviceGuard guard(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        set_alloc_stream(s);
        ILUB_CUDA(cudaEventRecord(h->ev[1], s));
        ILUB_CUDA(cudaMemcpyAsync(h->a.p, data, sizeof(double) * h->nnz, to_dev(where), s));
        if (h->prefactored)
            ILUB_CUDA(cudaMemcpyAsync(h->lu.p, h->a.p, sizeof(double) * h->nnz, cudaMemcpyDeviceToDevice, s));
        else
            factor_numeric(h, s);
        ILUB_CUDA(cudaEventRecord(h->ev[2], s));
        for (int k = 0; k < 4; ++k) pack_solve_values(h, k, s);
        ILUB_CUDA(cudaEventRecord(h->ev[3], s));
        read_status(h, s);
        float ms12 = 0, ms23 = 0;
        cudaEventElapsedTime(&ms12, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&ms23, h->ev[2], h->ev[3]);
        h->timings[0] = 0;
        h->timings[1] = ms12;
        h->timings[2] = ms23;
        h->timings[3] = ms12 + ms23;
    });
}

int ilub_destroy(ilub_handle* h) {
    return guarded([&] {
        if (!h) return;
        DeviceGuard guard(h->device);  // the caller's current device is left as it was
        cudaDeviceSynchronize();
        set_alloc_stream(kIdleStream);  // the device is idle: the handle's blocks can be reused on any stream
        delete h;
        set_alloc_stream(nullptr);
    });
}

int ilub_check_vector(const ilub_handle* h, int64_t len) {
    return guarded([&] {
        ILUB_REQUIRE(h, ILUB_ERR_INVALID, "null handle");
        ILUB_REQUIRE(len == h->n, ILUB_ERR_SIZE, "vector has wrong size for preconditioner!");
    });
}

int ilub_apply(ilub_handle* h, int transpose, double* x, int64_t ldx, int32_t nrhs, int where, void* stream) {
    return guarded([&] {
        ILUB_REQUIRE(h && x, ILUB_ERR_INVALID, "null argument");
        ILUB_REQUIRE(nrhs >= 1, ILUB_ERR_INVALID, "nrhs must be >= 1");
        ILUB_REQUIRE(ldx >= h->n, ILUB_ERR_SIZE, "vector has wrong size for preconditioner!");
        HandleLock lock(h->mu);
        OptionScope scope(h);
        DeviceGuard guard(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        set_alloc_stream(s);
        raise_deferred(h, s);
        int first, second;
        apply_slots(h, transpose, first, second);
        ensure_solve(h, first, s);
        ensure_solve(h, second, s);
        // pivoted factors (ilub_set_permutations): apply = scatter_out . T^-1 . gather_in, apply_trans = its transpose
        const int* pre = transpose ? (h->perm_out.n ? h->perm_out.p : nullptr) : (h->perm_in.n ? h->perm_in.p : nullptr);
        const int* post = transpose ? (h->perm_in.n ? h->perm_in.p : nullptr) : (h->perm_out.n ? h->perm_out.p : nullptr);
        if (where == ILUB_DEVICE) {
            if (pre) permute_gather(h, pre, x, ldx, nrhs, s);
            run_solve(h, first, x, ldx, nrhs, s);
            run_solve(h, second, x, ldx, nrhs, s);
            if (post) permute_scatter(h, post, x, ldx, nrhs, s);
            post_deferred(h, s);
        } else if (pre || post) {
            const size_t cnt = (size_t)h->n * nrhs;
            h->work.ensure(cnt);
            for (int c = 0; c < nrhs; ++c)
                ILUB_CUDA(cudaMemcpyAsync(h->work.p + (size_t)c * h->n, x + (size_t)c * ldx, sizeof(double) * h->n,
                                          cudaMemcpyHostToDevice, s));
            if (pre) permute_gather(h, pre, h->work.p, h->n, nrhs, s);
            run_solve(h, first, h->work.p, h->n, nrhs, s);
            run_solve(h, second, h->work.p, h->n, nrhs, s);
            if (post) permute_scatter(h, post, h->work.p, h->n, nrhs, s);
            for (int c = 0; c < nrhs; ++c)
                ILUB_CUDA(cudaMemcpyAsync(x + (size_t)c * ldx, h->work.p + (size_t)c * h->n, sizeof(double) * h->n,
                                          cudaMemcpyDeviceToHost, s));
            read_status(h, s);
        } else if (nrhs == 1 && options().zero_copy != 0 && systolic_out_of_place_ok(h, first) && systolic_out_of_place_ok(h, second) &&
                   device_view_of_host(x) != nullptr) {
            // Zero copy: x is pinned host memory the device can address.  The first solve reads the right-hand side
            // straight from it (PCIe reads trail the wavefront: no H2D copy ahead of the kernel), the second writes
            // the solution straight into it (posted PCIe writes behind the wavefront: no D2H copy after the kernel).
            double* xd = device_view_of_host(x);
            h->work.ensure((size_t)h->n);
            const int64_t zc = options().zero_copy;   // 1: both directions, 2: solution out only, 3: right-hand side in only
            if (zc == 2) ILUB_CUDA(cudaMemcpyAsync(h->work.p, x, sizeof(double) * h->n, cudaMemcpyHostToDevice, s));
            run_systolic_io(h, first, zc == 2 ? h->work.p : xd, h->work.p, s);
            run_systolic_io(h, second, h->work.p, zc == 3 ? h->work.p : xd, s);
            if (zc == 3) ILUB_CUDA(cudaMemcpyAsync(x, h->work.p, sizeof(double) * h->n, cudaMemcpyDeviceToHost, s));
            read_status(h, s);
        } else {
            const size_t cnt = (size_t)h->n * nrhs;
            h->work.ensure(cnt);
            for (int c = 0; c < nrhs; ++c)
                ILUB_CUDA(cudaMemcpyAsync(h->work.p + (size_t)c * h->n, x + (size_t)c * ldx, sizeof(double) * h->n,
                                          cudaMemcpyHostToDevice, s));
            run_solve(h, first, h->work.p, h->n, nrhs, s);
            run_solve(h, second, h->work.p, h->n, nrhs, s);
            for (int c = 0; c < nrhs; ++c)
                ILUB_CUDA(cudaMemcpyAsync(x + (size_t)c * ldx, h->work.p + (size_t)c * h->n, sizeof(double) * h->n,
                                          cudaMemcpyDeviceToHost, s));
            read_status(h, s);
        }
    });
}

int ilub_check_status(ilub_handle* h, void* stream) {
    return guarded([&] {
        ILUB_REQUIRE(h, ILUB_ERR_INVALID, "null handle");
        HandleLock lock(h->mu);
        OptionScope scope(h);
        DeviceGuard guard(h->device);
        read_status(h, static_cast<cudaStream_t>(stream));  // synchronises the stream; throws ILUB_ERR_WATCHDOG
        if (h->host_status) *(volatile int*)h->host_status = 0;
    });
}
int ilub_trim_cache(void) {
    return guarded([&] { pool_trim(); });
}

int ilub_total_nnz(const ilub_handle* h, int64_t* out) {
    return guarded([&] {
        ILUB_REQUIRE(h && out, ILUB_ERR_INVALID, "null argument");
        // preconditioner.h:164,208-209,246-247: left_nnz + right_nnz as stored (explicit unit diagonal included)
        *out = h->kind == KIND_ICHOL0 ? h->nnz_lower + h->n : h->nnz_lower + h->nnz_upper + 2 * (int64_t)h->n;
    });
}
int ilub_dim(const ilub_handle* h, int32_t* n) {
    return guarded([&] {
        ILUB_REQUIRE(h && n, ILUB_ERR_INVALID, "null argument");
        *n = h->n;
    });
}
int ilub_num_factors(const ilub_handle* h, int* count) {
    return guarded([&] {
        ILUB_REQUIRE(h && count, ILUB_ERR_INVALID, "null argument");
        *count = h->kind == KIND_ICHOL0 ? 1 : 2;
    });
}

// which factor arrays does reported factor `which` consist of?  mode as in extract_factor
static int factor_mode(const ilub_handle* h, int which) {
    if (h->kind == KIND_ICHOL0) return 2;
    // CSR input: L = unit lower (0), U = upper (1).  CSC input: ILU0.hpp:100-105 swaps them.
    return (h->is_csr != 0) == (which == ILUB_FACTOR_LEFT) ? 0 : 1;
}

int ilub_factor_dims(const ilub_handle* h, int which, int64_t* nnz, int* is_csr) {
    return guarded([&] {
        ILUB_REQUIRE(h && nnz && is_csr, ILUB_ERR_INVALID, "null argument");
        ILUB_REQUIRE(which == 0 || (which == 1 && h->kind == KIND_ILU0), ILUB_ERR_INVALID, "no such factor");
        const int mode = factor_mode(h, which);
        *nnz = (mode == 1 ? h->nnz_upper : h->nnz_lower) + h->n;
        *is_csr = h->kind == KIND_ICHOL0 ? 1 : h->is_csr;  // IChol.hpp:67-68: always ROW
    });
}

int ilub_get_factor(ilub_handle* h, int which, double* data, int32_t* indices, int32_t* indptr, int where,
                    void* stream) {
    return guarded([&] {
        ILUB_REQUIRE(h && data && indices && indptr, ILUB_ERR_INVALID, "null argument");
        ILUB_REQUIRE(which == 0 || (which == 1 && h->kind == KIND_ILU0), ILUB_ERR_INVALID, "no such factor");
        HandleLock lock(h->mu);
        OptionScope scope(h);
        DeviceGuard guard(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        set_alloc_stream(s);
        const int mode = factor_mode(h, which);
        const int64_t nnz = (mode == 1 ? h->nnz_upper : h->nnz_lower) + h->n;
        if (where == ILUB_DEVICE) {
            extract_factor(h, mode, data, indices, indptr, s);
        } else {
            DevBuf<double> d;
            DevBuf<int> i, p;
            d.alloc(nnz);
            i.alloc(nnz);
            p.alloc(h->n + 1);
            extract_factor(h, mode, d.p, i.p, p.p, s);
            ILUB_CUDA(cudaMemcpyAsync(data, d.p, sizeof(double) * nnz, cudaMemcpyDeviceToHost, s));
            ILUB_CUDA(cudaMemcpyAsync(indices, i.p, sizeof(int) * nnz, cudaMemcpyDeviceToHost, s));
            ILUB_CUDA(cudaMemcpyAsync(indptr, p.p, sizeof(int) * (h->n + 1), cudaMemcpyDeviceToHost, s));
            ILUB_CUDA(cudaStreamSynchronize(s));
        }
    });
}

// ---- persistence: the reference's raw `.isp` dump of a matrix_sparse<double> (sparse_implementation.h:3814-3860):
// int32 rows, cols, pointer_size, nnz; int32 orientation (ROW = 0, COLUMN = 1, declarations.h:83); then
// pointer[pointer_size], indices[nnz] (int32) and data[nnz] (double), nothing else
namespace {
struct IspHeader {
    int32_t rows, cols, pointer_size, nnz, orientation;
};
void isp_write_file(const std::string& path, int32_t rows, int32_t cols, int orientation, const double* data,
                    const int32_t* indices, const int32_t* indptr, int32_t nnz) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw Error{ILUB_ERR_INVALID, "cannot write " + path};
    const IspHeader hd = {rows, cols, (orientation == 0 ? rows : cols) + 1, nnz, orientation};
    bool ok = fwrite(&hd, sizeof(int32_t), 5, f) == 5;
    ok = ok && fwrite(indptr, sizeof(int32_t), (size_t)hd.pointer_size, f) == (size_t)hd.pointer_size;
    ok = ok && fwrite(indices, sizeof(int32_t), (size_t)nnz, f) == (size_t)nnz;
    ok = ok && fwrite(data, sizeof(double), (size_t)nnz, f) == (size_t)nnz;
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw Error{ILUB_ERR_INVALID, "error writing " + path};
}
IspHeader isp_read_header(FILE* f, const std::string& path) {
    IspHeader hd;
    if (fread(&hd, sizeof(int32_t), 5, f) != 5) throw Error{ILUB_ERR_INVALID, "error reading " + path + " (header)"};
    const bool sane = hd.rows >= 0 && hd.cols >= 0 && hd.nnz >= 0 && (hd.orientation == 0 || hd.orientation == 1) &&
                      hd.pointer_size == (hd.orientation == 0 ? hd.rows : hd.cols) + 1;
    if (!sane) throw Error{ILUB_ERR_INVALID, path + " is not an .isp matrix file (int32 indices, double values)"};
    return hd;
}
}  // namespace

int ilub_isp_write(const char* path, int32_t rows, int32_t cols, int is_csr, const double* data, const int32_t* indices,
                   const int32_t* indptr, int64_t nnz) {
    return guarded([&] {
        ILUB_REQUIRE(path && data && indices && indptr, ILUB_ERR_INVALID, "null argument");
        ILUB_REQUIRE(rows >= 0 && cols >= 0 && nnz >= 0 && nnz < ((int64_t)1 << 31), ILUB_ERR_SIZE, "bad matrix dimensions");
        isp_write_file(path, rows, cols, is_csr ? 0 : 1, data, indices, indptr, (int32_t)nnz);
    });
}
int ilub_isp_dims(const char* path, int32_t* rows, int32_t* cols, int64_t* nnz, int* is_csr) {
    return guarded([&] {
        ILUB_REQUIRE(path && rows && cols && nnz && is_csr, ILUB_ERR_INVALID, "null argument");
        FILE* f = fopen(path, "rb");
        if (!f) throw Error{ILUB_ERR_INVALID, std::string("cannot read ") + path};
        IspHeader hd;
        try {
            hd = isp_read_header(f, path);
        } catch (...) {
            fclose(f);
            throw;
        }
        fclose(f);
        *rows = hd.rows;
        *cols = hd.cols;
        *nnz = hd.nnz;
        *is_csr = hd.orientation == 0;
    });
}
int ilub_isp_read(const char* path, double* data, int32_t* indices, int32_t* indptr) {
    return guarded([&] {
        ILUB_REQUIRE(path && data && indices && indptr, ILUB_ERR_INVALID, "null argument");
        FILE* f = fopen(path, "rb");
        if (!f) throw Error{ILUB_ERR_INVALID, std::string("cannot read ") + path};
        bool ok = false;
        try {
            const IspHeader hd = isp_read_header(f, path);
            ok = fread(indptr, sizeof(int32_t), (size_t)hd.pointer_size, f) == (size_t)hd.pointer_size;
            ok = ok && fread(indices, sizeof(int32_t), (size_t)hd.nnz, f) == (size_t)hd.nnz;
            ok = ok && fread(data, sizeof(double), (size_t)hd.nnz, f) == (size_t)hd.nnz;
        } catch (...) {
            fclose(f);
            throw;
        }
        fclose(f);
        if (!ok) throw Error{ILUB_ERR_INVALID, std::string("error reading ") + path + " (truncated)"};
    });
}
int ilub_write_binary(ilub_handle* h, const char* filename, void* stream) {
    return guarded([&] {
        ILUB_REQUIRE(h && filename, ILUB_ERR_INVALID, "null argument");
        HandleLock lock(h->mu);
        const int nf = h->kind == KIND_ILU0 ? 2 : 1;
        for (int which = 0; which < nf; ++which) {
            int64_t nnz = 0;
            int is_csr = 1;
            if (ilub_factor_dims(h, which, &nnz, &is_csr) != ILUB_OK) throw Error{ILUB_ERR_INVALID, g_last_error};
            std::vector<double> d((size_t)nnz);
            std::vector<int32_t> i((size_t)nnz), p((size_t)h->n + 1);
            if (ilub_get_factor(h, which, d.data(), i.data(), p.data(), ILUB_HOST, stream) != ILUB_OK)
                throw Error{ILUB_ERR_CUDA, g_last_error};
            // file names of ILUCPreconditioner::write_binary (preconditioner_implementation.h:960-970)
            isp_write_file(std::string(filename) + (which == 0 ? "-L.isp" : "-R.isp"), h->n, h->n, is_csr ? 0 : 1, d.data(),
                           i.data(), p.data(), (int32_t)nnz);
        }
    });
}

int ilub_get_levels(ilub_handle* h, int which_solve, int32_t* nlevels, int32_t* level_of_row, int where,
                    void* stream) {
    return guarded([&] {
        ILUB_REQUIRE(h, ILUB_ERR_INVALID, "null handle");
        ILUB_REQUIRE(which_solve >= 0 && which_solve < 4, ILUB_ERR_INVALID, "bad which_solve");
        HandleLock lock(h->mu);
        OptionScope scope(h);
        DeviceGuard guard(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        set_alloc_stream(s);
        const int slot = slot_of(h, which_solve);
        ensure_solve(h, slot, s);
        if (nlevels) *nlevels = h->solve[slot].nlevels;
        if (level_of_row) {
            ILUB_CUDA(cudaMemcpyAsync(level_of_row, h->solve[slot].level.p, sizeof(int) * h->n, from_dev(where), s));
            if (where == ILUB_HOST) ILUB_CUDA(cudaStreamSynchronize(s));
        }
    });
}

int ilub_get_level_sets(ilub_handle* h, int which_solve, int32_t* level_ptr, int32_t* level_rows, int where,
                        void* stream) {
    return guarded([&] {
        ILUB_REQUIRE(h && level_ptr && level_rows, ILUB_ERR_INVALID, "null argument");
        ILUB_REQUIRE(which_solve >= 0 && which_solve < 4, ILUB_ERR_INVALID, "bad which_solve");
        HandleLock lock(h->mu);
        OptionScope scope(h);
        DeviceGuard guard(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        set_alloc_stream(s);
        const int slot = slot_of(h, which_solve);
        ensure_solve(h, slot, s);
        build_level_sets(h, slot, s);
        TriSolve& T = h->solve[slot];
        ILUB_CUDA(cudaMemcpyAsync(level_ptr, T.lvl_ptr.p, sizeof(int) * (T.nlevels + 1), from_dev(where), s));
        ILUB_CUDA(cudaMemcpyAsync(level_rows, T.lvl_rows.p, sizeof(int) * h->n, from_dev(where), s));
        if (where == ILUB_HOST) ILUB_CUDA(cudaStreamSynchronize(s));
    });
}

int ilub_spmv(ilub_handle* h, const double* x, double* y, int where, void* stream) {
    return guarded([&] {
        ILUB_REQUIRE(h && x && y, ILUB_ERR_INVALID, "null argument");
        ILUB_REQUIRE(!h->prefactored, ILUB_ERR_INVALID, "handle was created from factors: it holds no matrix");
        HandleLock lock(h->mu);
        OptionScope scope(h);
        DeviceGuard guard(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        set_alloc_stream(s);
        if (where == ILUB_DEVICE) {
            spmv(h, x, y, s);
        } else {
            DevBuf<double> dx, dy;
            dx.alloc(h->n);
            dy.alloc(h->n);
            ILUB_CUDA(cudaMemcpyAsync(dx.p, x, sizeof(double) * h->n, cudaMemcpyHostToDevice, s));
            spmv(h, dx.p, dy.p, s);
            ILUB_CUDA(cudaMemcpyAsync(y, dy.p, sizeof(double) * h->n, cudaMemcpyDeviceToHost, s));
            ILUB_CUDA(cudaStreamSynchronize(s));
        }
    });
}

int ilub_bicgstab(ilub_handle* h, const double* b, double* x, double rtol, double atol, int32_t max_iter,
                  int32_t* iters, double* rel_achieved, double* abs_achieved, int where, void* stream) {
    return guarded([&] {
        ILUB_REQUIRE(h && b && x && iters && rel_achieved && abs_achieved, ILUB_ERR_INVALID, "null argument");
        ILUB_REQUIRE(!h->prefactored, ILUB_ERR_INVALID, "handle was created from factors: it holds no matrix");
        HandleLock lock(h->mu);
        OptionScope scope(h);
        DeviceGuard guard(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        set_alloc_stream(s);
        int first, second;
        apply_slots(h, 0, first, second);
        ensure_solve(h, first, s);
        ensure_solve(h, second, s);
        int ok;
        if (where == ILUB_DEVICE) {
            ok = bicgstab(h, b, x, rtol, atol, max_iter, iters, rel_achieved, abs_achieved, s);
        } else {
            DevBuf<double> db, dx;
            db.alloc(h->n);
            dx.alloc(h->n);
            ILUB_CUDA(cudaMemcpyAsync(db.p, b, sizeof(double) * h->n, cudaMemcpyHostToDevice, s));
            ok = bicgstab(h, db.p, dx.p, rtol, atol, max_iter, iters, rel_achieved, abs_achieved, s);
            ILUB_CUDA(cudaMemcpyAsync(x, dx.p, sizeof(double) * h->n, cudaMemcpyDeviceToHost, s));
        }
        read_status(h, s);
        if (!ok) throw Error{ILUB_ERR_NOT_CONVERGED, "did not converge"};
    });
}

int ilub_pcg(ilub_handle* h, const double* b, double* x, double rtol, double atol, int32_t max_iter,
                  int32_t* iters, double* rel_achieved, double* abs_achieved, int where, void* stream) {
    return guarded([&] {
        ILUB_REQUIRE(h && b && x && iters && rel_achieved && abs_achieved, ILUB_ERR_INVALID, "null argument");
        ILUB_REQUIRE(!h->prefactored, ILUB_ERR_INVALID, "handle was created from factors: it holds no matrix");
        HandleLock lock(h->mu);
        OptionScope scope(h);
        DeviceGuard guard(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        set_alloc_stream(s);
        int first, second;
        apply_slots(h, 0, first, second);
        ensure_solve(h, first, s);
        ensure_solve(h, second, s);
        int ok;
        if (where == ILUB_DEVICE) {
            ok = pcg(h, b, x, rtol, atol, max_iter, iters, rel_achieved, abs_achieved, s);
        } else {
            DevBuf<double> db, dx;
            db.alloc(h->n);
            dx.alloc(h->n);
            ILUB_CUDA(cudaMemcpyAsync(db.p, b, sizeof(double) * h->n, cudaMemcpyHostToDevice, s));
            ok = pcg(h, db.p, dx.p, rtol, atol, max_iter, iters, rel_achieved, abs_achieved, s);
            ILUB_CUDA(cudaMemcpyAsync(x, dx.p, sizeof(double) * h->n, cudaMemcpyDeviceToHost, s));
        }
        read_status(h, s);
        if (!ok) throw Error{ILUB_ERR_NOT_CONVERGED, "did not converge"};
    });
}

int ilub_set_option(const char* key, int64_t value) {
    return guarded([&] {
        ILUB_REQUIRE(key, ILUB_ERR_INVALID, "null key");
        const int f = option_index(key);
        if (f < 0) throw Error{ILUB_ERR_INVALID, std::string("unknown option ") + key};
        set_default_option(f, value);
    });
}
int ilub_get_option(const char* key, int64_t* value) {
    return guarded([&] {
        ILUB_REQUIRE(key && value, ILUB_ERR_INVALID, "null argument");
        const int f = option_index(key);
        if (f < 0) throw Error{ILUB_ERR_INVALID, std::string("unknown option ") + key};
        OptionScope none(nullptr);
        *value = options().*(option_fields(nullptr)[f].member);
    });
}
int ilub_stage_option(const char* key, int64_t value) {
    return guarded([&] {
        ILUB_REQUIRE(key, ILUB_ERR_INVALID, "null key");
        const int f = option_index(key);
        if (f < 0) throw Error{ILUB_ERR_INVALID, std::string("unknown option ") + key};
        tl_staged_options.emplace_back(f, value);
    });
}
int ilub_set_handle_option(ilub_handle* h, const char* key, int64_t value) {
    return guarded([&] {
        ILUB_REQUIRE(h && key, ILUB_ERR_INVALID, "null argument");
        const int f = option_index(key);
        if (f < 0) throw Error{ILUB_ERR_INVALID, std::string("unknown option ") + key};
        HandleLock lock(h->mu);
        for (auto& ov : h->opt_overrides)
            if (ov.first == f) {
                ov.second = value;
                return;
            }
        h->opt_overrides.emplace_back(f, value);
    });
}
int ilub_get_handle_option(ilub_handle* h, const char* key, int64_t* value) {
    return guarded([&] {
        ILUB_REQUIRE(h && key && value, ILUB_ERR_INVALID, "null argument");
        const int f = option_index(key);
        if (f < 0) throw Error{ILUB_ERR_INVALID, std::string("unknown option ") + key};
        HandleLock lock(h->mu);
        OptionScope scope(h);
        *value = options().*(option_fields(nullptr)[f].member);
    });
}

/* debug (not part of the supported ABI): per-task block time stamps of the line-systolic solve (option trace) */
ILUB_API int ilub_debug_sys_trace(ilub_handle* h, int which_solve, int32_t* dims /*[8]*/, long long* trace) {
    return guarded([&] {
        ILUB_REQUIRE(h && dims, ILUB_ERR_INVALID, "null argument");
        DeviceGuard guard(h->device);
        TriSolve& T = h->solve[slot_of(h, which_solve)];
        ILUB_REQUIRE(T.built && T.sys.built, ILUB_ERR_INVALID, "no line-systolic solve");
        const SysSolve& Y = T.sys;
        if (Y.tile) {  // tile-systolic: per task the start time of every stage, then the end time
            dims[0] = Y.ntasks;
            dims[1] = 3 * (Y.nsteps_pad / 8 + 1);  // per task: stage start times + end time, data-wait cycles, gate-wait cycles
            dims[2] = Y.nty;
            dims[3] = Y.ntz;
            dims[4] = Y.twy;
            dims[5] = Y.twz;
            dims[6] = 0;
            dims[7] = Y.tnst * 100 + Y.tring;
            ILUB_CUDA(cudaDeviceSynchronize());
            const size_t cnt = (size_t)dims[0] * dims[1];
            if (trace && T.trace.n >= cnt) ILUB_CUDA(cudaMemcpy(trace, T.trace.p, sizeof(long long) * cnt, cudaMemcpyDeviceToHost));
            return;
        }
        dims[0] = Y.ntasks;
        dims[1] = 3 * (Y.nsteps_pad / 8 + 1);  // per task: time, failed polls, helper/stage wait cycles
        dims[2] = Y.nyb;
        dims[3] = Y.nplanes;
        dims[4] = Y.pyb;
        dims[5] = Y.pz;
        dims[6] = Y.csize;
        dims[7] = Y.nstages * 100 + Y.rz;
        ILUB_CUDA(cudaDeviceSynchronize());
        const size_t cnt = (size_t)dims[0] * dims[1];
        if (trace && T.trace.n >= cnt) ILUB_CUDA(cudaMemcpy(trace, T.trace.p, sizeof(long long) * cnt, cudaMemcpyDeviceToHost));
    });
}

/* debug (not part of the supported ABI): chunk -> row map and per-chunk clock stamps of one solve */
ILUB_API int ilub_debug_chunks(ilub_handle* h, int which_solve, int32_t* nchunks, int32_t* rows, long long* trace) {
    return guarded([&] {
        ILUB_REQUIRE(h && nchunks, ILUB_ERR_INVALID, "null argument");
        DeviceGuard guard(h->device);
        TriSolve& T = h->solve[slot_of(h, which_solve)];
        ILUB_REQUIRE(T.built, ILUB_ERR_INVALID, "solve not built");
        *nchunks = T.nchunks;
        ILUB_CUDA(cudaDeviceSynchronize());
        if (rows) ILUB_CUDA(cudaMemcpy(rows, T.rows.p, sizeof(int) * (size_t)T.nchunks * 32, cudaMemcpyDeviceToHost));
        if (trace && T.trace.n >= (size_t)T.nchunks * 4)
            ILUB_CUDA(cudaMemcpy(trace, T.trace.p, sizeof(long long) * (size_t)T.nchunks * 4, cudaMemcpyDeviceToHost));
    });
}

int ilub_get_timings(const ilub_handle* h, double out_ms[4]) {
    return guarded([&] {
        ILUB_REQUIRE(h && out_ms, ILUB_ERR_INVALID, "null argument");
        for (int i = 0; i < 4; ++i) out_ms[i] = h->timings[i];
    });
}
int ilub_get_stats(const ilub_handle* h, int64_t out[16]) {
    return guarded([&] {
        ILUB_REQUIRE(h && out, ILUB_ERR_INVALID, "null argument");
        HandleLock lock(const_cast<ilub_handle*>(h)->mu);
        OptionScope scope(const_cast<ilub_handle*>(h));
        out[0] = h->n;
        out[1] = h->nnz;
        out[2] = h->kind == KIND_ICHOL0 ? h->nnz_lower + h->n : h->nnz_lower + h->nnz_upper + 2 * (int64_t)h->n;
        for (int k = 0; k < 4; ++k) {
            const int slot = slot_of(h, k);
            out[3 + k] = h->solve[slot].built ? h->solve[slot].nlevels : 0;
        }
        out[7] = h->launches;
        const TriSolve& T = h->solve[slot_of(h, 0)];
        out[8] = T.nparts;
        out[9] = T.part_mode;
        out[10] = T.grid_y;
        out[11] = T.grid_z;
        out[12] = T.nchunks;
        out[13] = T.nslabs;
        out[14] = T.cfg_warps;
        out[15] = (int64_t)T.cfg_smem;
        if (T.sys.built && options().solve_algo == 0) {  // line-systolic path: tasks instead of chunks
            const SysSolve& Y = T.sys;
            out[8] = (int64_t)Y.gyb * Y.gz;
            out[9] = 4;
            out[10] = Y.gyb;
            out[11] = Y.gz;
            out[12] = Y.ntasks * 100 + Y.csize;  // tasks, cluster size in the last two digits
            out[13] = Y.nsteps_pad;
            out[14] = Y.pyb * Y.pz + 2;
            out[15] = (int64_t)Y.smem;
            if (Y.tile) {  // tile-systolic: CTA grid of twy x twz tiles each
                out[8] = (int64_t)Y.tgy * Y.tgz;
                out[9] = 7;
                out[10] = Y.tgy;
                out[11] = Y.tgz;
                out[14] = Y.twy * Y.twz;
            }
        }
    });
}

}  // extern "C"
