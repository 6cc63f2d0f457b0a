// dev: latency microbenchmarks behind the cost model of the line-systolic solve.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench scripts/ubench.cu ; run on the GPU box.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <climits>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld_s128(uint32_t a, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.cta.shared.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}" : "=l"(lo), "=l"(hi) : "r"(a) : "memory");
}
__device__ __forceinline__ void st_s128(uint32_t a, double x, unsigned long long tag) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.cta.shared.b128 [%0], q;\n\t}" ::"r"(a), "l"(__double_as_longlong(x)), "l"(tag) : "memory");
}
__device__ __forceinline__ void st_gen128(void* p, double x, unsigned long long tag) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.gpu.b128 [%0], q;\n\t}" ::"l"(p), "l"(__double_as_longlong(x)), "l"(tag) : "memory");
}
__device__ __forceinline__ void ld_g128(const void* p, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_cluster128(uint32_t a, double x, unsigned long long tag) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.cluster.shared::cluster.b128 [%0], q;\n\t}" ::"r"(a), "l"(__double_as_longlong(x)), "l"(tag) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}

// ---- 1: dependent DFMA chain
__global__ void k_dfma(double* out, long long* cyc, int n) {
    double a = out[0], b = out[1], x = out[2];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = fma(a, x, b);
    long long t1 = clock64();
    out[3 + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// ---- 2: shfl.up (64-bit) + 2 DFMA per step
__global__ void k_shfl(double* out, long long* cyc, int n) {
    double a = out[0], b = out[1], x = out[2] + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; ++i) {
        double y = __shfl_up_sync(0xffffffffu, x, 1);
        double acc = fma(a, x, b);
        x = fma(a, y, acc);
    }
    long long t1 = clock64();
    out[3 + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// ---- 3: step-like: shfl + fma + poll of an always-ready smem entry + vote + sts + 16-byte ring store
// mode bit0: poll+vote, bit1: window store, bit2: ring store (shared), bit3: generic ring store, bit4: spinning neighbour warps
__global__ void k_step(double* out, long long* cyc, int n, int mode) {
    __shared__ __align__(16) ulonglong2 ring[16 * 32];
    __shared__ double win[32 * 65];
    __shared__ int prog[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) ring[i] = make_ulonglong2(0ull, (unsigned long long)INT_MAX);
    if (threadIdx.x < 8) prog[threadIdx.x] = 0;
    __syncthreads();
    if (warp >= 4) {  // spinning neighbours (same SMSP as warp - 4)
        if (mode & 16) {
            while (*(volatile int*)&prog[warp - 4] == 0) {}
        }
        return;
    }
    double a = out[0], b = out[1], x = out[2] + lane;
    const uint32_t rs = smem_u32(ring) + lane * 16;
    const uint32_t ws = smem_u32(win) + lane * 65 * 8;
    ulonglong2* rg = ring + lane;
    long long t0 = clock64();
    for (int i0 = 0; i0 < n; i0 += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u;
            unsigned long long vz = 0, tz = INT_MAX;
            if (mode & 1) ld_s128(rs + (((i + 8) & 15) << 9), vz, tz);
            double y = __shfl_up_sync(0xffffffffu, x, 1);
            double acc = fma(a, x, b);
            acc = fma(a, y, acc);
            double xn;
            while (true) {
                xn = fma(a, __longlong_as_double((long long)vz), acc);
                if (!(mode & 1) || __all_sync(0xffffffffu, (int)(unsigned)tz >= i + 1)) break;
                ld_s128(rs + (((i + 8) & 15) << 9), vz, tz);
            }
            x = xn;
            if (mode & 2) asm volatile("st.shared.f64 [%0], %1;" ::"r"(ws + ((i & 63) << 3)), "d"(x));
            if (mode & 4) st_s128(rs + ((i & 7) << 9), x, (unsigned long long)INT_MAX);
            if (mode & 8) st_gen128(rg + ((i & 7) << 5), x, (unsigned long long)INT_MAX);
        }
    }
    long long t1 = clock64();
    out[3 + threadIdx.x] = x;
    if (lane == 0) {
        cyc[warp] = t1 - t0;
        *(volatile int*)&prog[warp] = 1;
    }
}
// ---- 4: ping-pong through shared memory between warp 0 and warp `other` of one CTA
__global__ void k_pp_smem(long long* cyc, int n, int other, int generic) {
    __shared__ __align__(16) ulonglong2 box[2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 64) box[threadIdx.x >> 5][threadIdx.x & 31] = make_ulonglong2(0, 0);
    __syncthreads();
    if (warp != 0 && warp != other) return;
    const int me = warp == 0 ? 0 : 1;
    const uint32_t mine = smem_u32(&box[me][lane]), theirs = smem_u32(&box[1 - me][lane]);
    ulonglong2* theirs_g = &box[1 - me][lane];
    long long t0 = clock64();
    double x = 1.0;
    for (int i = 1; i <= n; ++i) {
        if (me == 0) {
            if (generic) st_gen128(theirs_g, x, i); else st_s128(theirs, x, i);
        }
        unsigned long long v, t;
        do {
            ld_s128(mine, v, t);
        } while (!__all_sync(0xffffffffu, (int)t >= i));
        x = __longlong_as_double((long long)v) + 1.0;
        if (me == 1) {
            if (generic) st_gen128(theirs_g, x, i); else st_s128(theirs, x, i);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// ---- 5: ping-pong through distributed shared memory (cluster of 2)
__global__ void __cluster_dims__(2, 1, 1) k_pp_dsmem(long long* cyc, int n) {
    __shared__ __align__(16) ulonglong2 box[32];
    const int lane = threadIdx.x & 31;
    cg::cluster_group cl = cg::this_cluster();
    const int me = cl.block_rank();
    box[lane] = make_ulonglong2(0, 0);
    cl.sync();
    const uint32_t mine = smem_u32(&box[lane]);
    const uint32_t theirs = mapa(mine, 1 - me);
    long long t0 = clock64();
    double x = 1.0;
    for (int i = 1; i <= n; ++i) {
        if (me == 0) st_cluster128(theirs, x, i);
        unsigned long long v, t;
        do {
            ld_s128(mine, v, t);
        } while (!__all_sync(0xffffffffu, (int)t >= i));
        x = __longlong_as_double((long long)v) + 1.0;
        if (me == 1) st_cluster128(theirs, x, i);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && me == 0) cyc[0] = t1 - t0;
    cl.sync();
}
// ---- 6: ping-pong through global memory between CTA 0 and CTA `other`
__global__ void k_pp_gmem(ulonglong2* box, long long* cyc, int n, int other) {
    if (blockIdx.x != 0 && blockIdx.x != other) return;
    const int lane = threadIdx.x & 31;
    const int me = blockIdx.x == 0 ? 0 : 1;
    ulonglong2* mine = box + me * 32 + lane;
    ulonglong2* theirs = box + (1 - me) * 32 + lane;
    long long t0 = clock64();
    double x = 1.0;
    for (int i = 1; i <= n; ++i) {
        if (me == 0) st_gen128(theirs, x, i);
        unsigned long long v, t;
        do {
            ld_g128(mine, v, t);
        } while (!__all_sync(0xffffffffu, (int)t >= i));
        x = __longlong_as_double((long long)v) + 1.0;
        if (me == 1) st_gen128(theirs, x, i);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && me == 0) cyc[0] = t1 - t0;
}
// ---- 7: chain of W warps in one CTA, each step passes {x, tag} to the next warp (systolic z chain), n steps:
// time of the last warp per step = rate; (end of last - end of first) / (W - 1) = hop lag
__global__ void k_chain(double* out, long long* cyc, int n, int W) {
    __shared__ __align__(16) ulonglong2 ring[8][8 * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * 8 * 32; i += blockDim.x) (&ring[0][0])[i] = make_ulonglong2(0ull, warp == 0 ? 0ull : 0ull);
    __shared__ int prog[8];
    if (threadIdx.x < 8) prog[threadIdx.x] = 0;
    __syncthreads();
    if (warp >= W) return;
    double a = out[0], b = out[1], x = out[2] + lane;
    const uint32_t in = smem_u32(&ring[warp][lane]);
    const uint32_t outp = smem_u32(&ring[(warp + 1) & 7][lane]);
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        unsigned long long vz = 0, tz = INT_MAX;
        if (warp > 0) ld_s128(in + ((i & 7) << 9), vz, tz);
        double y = __shfl_up_sync(0xffffffffu, x, 1);
        double acc = fma(a, x, b);
        acc = fma(a, y, acc);
        double xn;
        while (true) {
            xn = fma(a, __longlong_as_double((long long)vz), acc);
            if (__all_sync(0xffffffffu, (int)(unsigned)tz >= i + 1)) break;
            ld_s128(in + ((i & 7) << 9), vz, tz);
        }
        x = xn;
        if (warp + 1 < W) {
            // flow control: consumer at most 12 steps behind
            if ((i & 3) == 0) while (*(volatile int*)&prog[warp + 1] < i - 4) {}
            st_s128(outp + ((i & 7) << 9), x, (unsigned long long)(i + 1));
        }
        if ((i & 3) == 3 && lane == 0) *(volatile int*)&prog[warp] = i + 1;
    }
    long long t1 = clock64();
    out[3 + threadIdx.x] = x;
    if (lane == 0) { cyc[warp] = t1 - t0; cyc[8 + warp] = t1; }
}

int main() {
    double* out; long long* cyc; ulonglong2* box;
    CK(cudaMalloc(&out, 8 * 4096)); CK(cudaMalloc(&cyc, 8 * 64)); CK(cudaMalloc(&box, 16 * 64));
    double h[3] = {0.999, 0.001, 1.0};
    CK(cudaMemcpy(out, h, 24, cudaMemcpyHostToDevice));
    long long hc[16];
    const int n = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        k_dfma<<<1, 32>>>(out, cyc, n); CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
        if (rep) printf("dfma chain: %.1f cyc/op\n", (double)hc[0] / n);
        k_shfl<<<1, 32>>>(out, cyc, n); CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
        if (rep) printf("shfl64+2dfma step: %.1f cyc/step\n", (double)hc[0] / n);
        for (int mode : {0, 1, 3, 7, 11, 15, 31}) {
            for (int nw : {1, 4}) {
                k_step<<<1, (mode & 16) ? 256 : nw * 32>>>(out, cyc, n, mode); CK(cudaMemcpy(hc, cyc, 32, cudaMemcpyDeviceToHost));
                if (rep) printf("step mode %2d warps %d: %.1f cyc/step\n", mode, nw, (double)hc[0] / n);
            }
        }
        for (int other : {1, 4}) for (int g : {0, 1}) {
            k_pp_smem<<<1, 256>>>(cyc, n, other, g); CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
            if (rep) printf("smem ping-pong warp0<->warp%d generic=%d: %.1f cyc/hop\n", other, g, (double)hc[0] / n / 2);
        }
        k_pp_dsmem<<<2, 32>>>(cyc, n); CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
        if (rep) printf("dsmem ping-pong: %.1f cyc/hop\n", (double)hc[0] / n / 2);
        for (int other : {1, 2, 75}) {
            CK(cudaMemset(box, 0, 16 * 64));
            k_pp_gmem<<<148, 32>>>(box, cyc, n, other); CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
            if (rep) printf("gmem ping-pong cta0<->cta%d: %.1f cyc/hop\n", other, (double)hc[0] / n / 2);
        }
        for (int W : {1, 2, 4, 8}) {
            k_chain<<<1, 256>>>(out, cyc, n, W); CK(cudaMemcpy(hc, cyc, 128, cudaMemcpyDeviceToHost));
            if (rep) printf("chain W=%d: first warp %.1f cyc/step, last warp %.1f cyc/step, end lag per hop %.1f cyc\n", W, (double)hc[0] / n,
                            (double)hc[W - 1] / n, W > 1 ? (double)(hc[8 + W - 1] - hc[8]) / (W - 1) : 0.0);
        }
    }
    CK(cudaDeviceSynchronize());
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock %d kHz\n", clk);
    return 0;
}
