// Caching device allocator behind DevBuf (ilub_common.cuh).  The symbolic phase allocates and frees dozens of
// temporaries; plain cudaMalloc / cudaFree are device-synchronising and cost up to milliseconds each, which made
// the phase 5-25x slower than its kernels.  Freed blocks are kept in size-ordered free lists, tagged with the stream
// the freeing thread works on, and only handed out again for work on that stream: stream order then keeps a reused
// block safe without any device synchronisation.
#include <map>
#include <mutex>
#include <unordered_map>

#include "ilub_common.cuh"

namespace ilub {
namespace {
struct PoolBlock {
    void* p;
    size_t bytes;
    int device;
    cudaStream_t stream;
};
struct Pool {
    std::mutex mu;
    std::multimap<size_t, PoolBlock> free_blocks;           // by size
    std::unordered_map<void*, PoolBlock> live;              // handed out
    size_t cached = 0;
    static constexpr size_t kMaxCached = (size_t)24 << 30;  // beyond this, freed blocks go back to the driver
};
Pool& pool() {
    static Pool* p = new Pool();  // never destroyed: handles may be released during static destruction
    return *p;
}
thread_local cudaStream_t tls_alloc_stream = nullptr;
size_t round_size(size_t b) {
    if (b < ((size_t)1 << 20)) return (b + 511) & ~(size_t)511;
    return (b + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
}
}  // namespace

void set_alloc_stream(cudaStream_t s) { tls_alloc_stream = s; }

void pool_trim() noexcept {
    Pool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    int cur = -1;
    cudaGetDevice(&cur);
    for (auto& kv : P.free_blocks) {
        cudaSetDevice(kv.second.device);
        cudaFree(kv.second.p);
    }
    if (cur >= 0) cudaSetDevice(cur);
    P.free_blocks.clear();
    P.cached = 0;
}

void* pool_alloc(size_t bytes) {
    Pool& P = pool();
    const size_t want = round_size(bytes);
    int dev = 0;
    ILUB_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(P.mu);
        for (auto it = P.free_blocks.lower_bound(want); it != P.free_blocks.end() && it->first <= want + want / 4 + 4096; ++it) {
            if (it->second.device == dev && (it->second.stream == tls_alloc_stream || it->second.stream == kIdleStream)) {
                PoolBlock b = it->second;
                P.free_blocks.erase(it);
                P.cached -= b.bytes;
                P.live[b.p] = b;
                return b.p;
            }
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {  // give the cache back and try once more
        cudaGetLastError();
        pool_trim();
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess)
        throw Error{ILUB_ERR_ALLOC, "cudaMalloc of " + std::to_string(want) + " bytes: " + cudaGetErrorString(e)};
    std::lock_guard<std::mutex> lk(P.mu);
    P.live[p] = PoolBlock{p, want, dev, tls_alloc_stream};
    return p;
}

void pool_free(void* p) noexcept {
    if (!p) return;
    Pool& P = pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.live.find(p);
        if (it != P.live.end()) {
            PoolBlock b = it->second;
            P.live.erase(it);
            if (P.cached + b.bytes <= Pool::kMaxCached) {
                b.stream = tls_alloc_stream;  // later users on this stream are ordered behind whatever still reads it
                P.free_blocks.emplace(b.bytes, b);
                P.cached += b.bytes;
                return;
            }
        }
    }
    cudaFree(p);
}

}  // namespace ilub
