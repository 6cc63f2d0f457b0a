// dev: replica of the compute-warp step loop of k_trsv_systolic (fast arithmetic, z-coupled) with features that can be
// switched off at compile time, to find what a step costs on an otherwise idle SM.  One warp, all inputs "always ready".
#include <cstdio>
#include <cstdint>
#include <climits>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr unsigned FULL = 0xffffffffu;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld_s128(uint32_t a, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.cta.shared.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}" : "=l"(lo), "=l"(hi) : "r"(a));
}
__device__ __forceinline__ void st_gen128(void* p, double x, unsigned long long tag) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.gpu.b128 [%0], q;\n\t}" ::"l"(p), "l"(__double_as_longlong(x)), "l"(tag) : "memory");
}
__device__ __forceinline__ void st_s128(uint32_t a, double x, unsigned long long tag) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.cta.shared.b128 [%0], q;\n\t}" ::"r"(a), "l"(__double_as_longlong(x)), "l"(tag) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void lds_2f64(uint32_t a, double& v0, double& v1) { asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a)); }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ int ld_prog(const int* p) { return *(const volatile int*)p; }

enum { F_NOINL = 1024, F_ZPOLL = 1, F_YPOLL = 2, F_ZOUT = 4, F_YOUT = 8, F_WIN = 16, F_COEF = 32, F_FLOW = 64, F_SLOW = 128, F_VOTE = 256, F_ZSHARED = 512 };

__device__ __noinline__ ulonglong2 slow_poll(uint32_t zaddr, uint32_t yaddr, int tag, int flags, int* status, unsigned spin_iters) {
    unsigned long long vz = 0ull, tz = INT_MAX, vy = 0ull, ty = INT_MAX;
    unsigned spins = 0;
    do {
        if ((++spins & 0xfffu) == 0) { if (spins >= spin_iters || *(volatile int*)status != 0) break; }
        if (flags & 1) ld_s128(zaddr, vz, tz);
        if (flags & 2) ld_s128(yaddr, vy, ty);
    } while (!__all_sync(FULL, (int)(unsigned)tz >= tag && (int)(unsigned)ty >= tag + 32));
    return make_ulonglong2(vz, vy);
}

template <int F>
__global__ void __launch_bounds__(128, 1) k_step(double* out, long long* cyc, int nblocks, int spin_iters, int bg, const double* gsrc, double* gdst) {
    extern __shared__ __align__(128) unsigned char smem[];
    // layout: 3 stages of 8 KB | window 32 x 65 doubles | z ring 16 x 512 | y ring 64 x 16 | out ring 16 x 512 | ctl
    const int lane = threadIdx.x & 31;
    unsigned char* stage = smem;
    unsigned char* win = smem + 3 * 8192;
    unsigned char* zr = win + 32 * 65 * 8;
    unsigned char* yr = zr + 16 * 512;
    unsigned char* zo = yr + 64 * 16;
    int* ctl = (int*)(zo + 16 * 512);
    for (int i = threadIdx.x; i < 3 * 1024; i += blockDim.x) ((double*)stage)[i] = 1e-3 * (i % 7);
    for (int i = threadIdx.x; i < 32 * 65; i += blockDim.x) ((double*)win)[i] = 1.0;
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) ((ulonglong2*)zr)[i] = make_ulonglong2(0ull, (unsigned long long)INT_MAX);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) ((ulonglong2*)yr)[i] = make_ulonglong2(0ull, (unsigned long long)INT_MAX);
    if (threadIdx.x < 8) ctl[threadIdx.x] = threadIdx.x == 5 ? 0 : INT_MAX;
    __syncthreads();
    if (threadIdx.x >= 32) {
        const int w = threadIdx.x >> 5;
        volatile int* done = (volatile int*)&ctl[5];
        unsigned char* scratch = zo + 16 * 512 + 1024;  // 16 KB of scratch behind the rings
        if (w == 1 && (bg & 1)) {  // TMA bulk copies into scratch, 8 KB each, back to back
            uint64_t* bar = (uint64_t*)(scratch + 16384);
            if (lane == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                unsigned ph = 0;
                long long i = 0;
                while (*done == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(8192) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(scratch)), "l"(gsrc + (i & 1023) * 1024), "r"(8192), "r"(smem_u32(bar)) : "memory");
                    unsigned ok = 0;
                    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(ph) : "memory");
                    ph ^= 1; ++i;
                }
            }
        }
        if (w == 1 && (bg & 8)) {  // coalesced LDG / STG streaming like the rectangle mover
            long long i = 0;
            while (*done == 0) {
                double v[8];
                for (int q = 0; q < 8; ++q) v[q] = __ldcs(gsrc + ((i * 8 + q) & 8191) * 32 + lane);
                for (int q = 0; q < 8; ++q) __stcs(gdst + ((i * 8 + q) & 8191) * 32 + lane, v[q]);
                ++i;
            }
        }
        if (w == 2 && (bg & 2)) {  // shared-memory load/store traffic
            double a = 0;
            while (*done == 0) {
                for (int q = 0; q < 8; ++q) a += lds_f64(smem_u32(scratch) + 8192 + ((lane * 8 + q * 256) & 8191));
                for (int q = 0; q < 8; ++q) sts_f64(smem_u32(scratch) + 8192 + ((lane * 8 + q * 256) & 8191), a);
            }
            out[100 + lane] = a;
        }
        if (w == 3 && (bg & 4)) {  // global polling like a mailbox fetcher
            unsigned long long lo, hi, acc = 0;
            while (*done == 0) {
                for (int q = 0; q < 4; ++q) {
                    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}" : "=l"(lo), "=l"(hi) : "l"(gsrc + q * 64 + lane * 2));
                    acc += lo + hi;
                }
            }
            out[200 + lane] = (double)acc;
        }
        return;
    }
    const uint32_t stage_s = smem_u32(stage) + lane * 16, xw_lane = smem_u32(win) + lane * 65 * 8;
    const uint32_t zin_s = smem_u32(zr) + lane * 16, yin_s = smem_u32(yr);
    unsigned char* zdst0 = zo + lane * 16;
    unsigned char* ydst0 = zo;  // lane 31 only: shares the sink with the z hand-offs
    const int* zcons = &ctl[0];
    const int* ycons = &ctl[1];
    const bool has_zout = out[5] == 0.0, ypred = has_zout && lane == 31;
    const unsigned long long ehi = 7ull << 32;
    unsigned spins = 0;
    int* status = (int*)&ctl[4];
    double xprev = 0.0, xy = 0.0, acc1 = 0.0, ycur = 0.0;
    double a_cz, a_cy, a_cx, a_dv, a_bb, n_cz, n_cy, n_cx, n_dv, n_bb;
    unsigned long long vz = 0ull, tz = INT_MAX, vy = 0ull, ty = INT_MAX;
    lds_2f64(stage_s, a_cz, a_cy); lds_2f64(stage_s + 512, a_cx, a_dv); a_bb = lds_f64(xw_lane);
    lds_2f64(stage_s + 1024, n_cz, n_cy); lds_2f64(stage_s + 1024 + 512, n_cx, n_dv); n_bb = lds_f64(xw_lane + 8);
    acc1 = a_bb * a_dv;
    bool ok = true;
    int zseen = 0, yseen = 0, slot = 0;
    long long t0 = clock64();
    for (int B = 0; B < nblocks; ++B) {
        const int slot_n = slot + 1 == 3 ? 0 : slot + 1;
        const uint32_t st = stage_s + slot * 8192, stn = stage_s + slot_n * 8192;
        const uint32_t wb = xw_lane + ((B & 7) << 6), wbn = xw_lane + (((B + 1) & 7) << 6);
        const uint32_t zb = zin_s + (((unsigned)(B * 8) & 15u) << 9), zbn = zin_s + (((unsigned)((B + 1) * 8) & 15u) << 9);
        const uint32_t yb = yin_s + (((B + 4) & 7) << 7), ybn = yin_s + (((B + 5) & 7) << 7);
        unsigned char* zdst = zdst0 + (((unsigned)(B * 8) & 15u) << 9);
        unsigned char* ydst = ydst0 + ((B & 7) << 7);
        asm volatile("" : "+l"(zdst), "+l"(ydst));
        const int s0 = B * 8;
        int zpre = 0, ypre = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int tag = s0 + u + 1;
            if ((F & F_FLOW) && (u & 3) == 0) {
                const int h = 2 * B + (u >> 2);
                while (zseen < h - 3) zseen = ld_prog(zcons);
                while (yseen < h - 23) yseen = ld_prog(ycons);
            }
            if ((F & F_NOINL) && __builtin_expect(!ok, 0)) {
                const ulonglong2 r = slow_poll(zb + u * 512, yb + u * 16, tag, 3, status, (unsigned)spin_iters);
                vz = r.x, vy = r.y;
            }
            if ((F & F_SLOW) && !ok) {
                do {
                    if ((++spins & 0xfffu) == 0) { if (spins >= (unsigned)spin_iters || *(volatile int*)status != 0) return; }
                    if (F & F_ZPOLL) ld_s128(zb + u * 512, vz, tz);
                    if (F & F_YPOLL) ld_s128(yb + u * 16, vy, ty);
                } while (!__all_sync(FULL, (int)(unsigned)tz >= tag && (int)(unsigned)ty >= tag + 32));
            }
            unsigned long long vzn = 0ull, tzn = INT_MAX, vyn = 0ull, tyn = INT_MAX;
            if (u + 1 < 8) {
                if (F & F_ZPOLL) ld_s128(zb + (u + 1) * 512, vzn, tzn);
                if (F & F_YPOLL) ld_s128(yb + (u + 1) * 16, vyn, tyn);
            } else {
                if (F & F_ZPOLL) ld_s128(zbn, vzn, tzn);
                if (F & F_YPOLL) ld_s128(ybn, vyn, tyn);
            }
            if (lane == 0) xy = ycur;
            const double acc2 = fma(-a_cy, xy, acc1);
            const double xnew = fma(-a_cz, __longlong_as_double((long long)vz), acc2);
            const unsigned long long tag64 = ehi | (unsigned long long)(unsigned)tag;
            if (F & F_ZOUT) { if (has_zout) { if (F & F_ZSHARED) st_s128(smem_u32(zdst) + u * 512, xnew, tag64); else st_gen128(zdst + u * 512, xnew, tag64); } }
            xy = __shfl_up_sync(FULL, xnew, 1);
            acc1 = fma(-n_cx, xnew, n_bb * n_dv);
            xprev = xnew;
            ycur = __longlong_as_double((long long)vy);
            if (F & F_WIN) sts_f64(wb + u * 8, xnew);
            if (F & F_YOUT) { if (ypred) st_gen128(ydst + u * 16, xnew, tag64); }
            if ((F & F_FLOW) && (u & 3) == 3 && lane == 0) *(volatile int*)&ctl[2] = 2 * B + (u >> 2) + 1;
            vz = vzn, tz = tzn, vy = vyn, ty = tyn;
            if (F & F_VOTE) ok = __all_sync(FULL, (int)(unsigned)tz >= tag + 1 && (int)(unsigned)ty >= tag + 33);
            a_cz = n_cz, a_cy = n_cy, a_cx = n_cx, a_dv = n_dv, a_bb = n_bb;
            if (F & F_COEF) {
                if (u + 2 < 8) {
                    lds_2f64(st + (u + 2) * 1024, n_cz, n_cy); lds_2f64(st + (u + 2) * 1024 + 512, n_cx, n_dv); n_bb = lds_f64(wb + (u + 2) * 8);
                } else {
                    lds_2f64(stn + (u + 2 - 8) * 1024, n_cz, n_cy); lds_2f64(stn + (u + 2 - 8) * 1024 + 512, n_cx, n_dv); n_bb = lds_f64(wbn + (u + 2 - 8) * 8);
                }
            }
            if ((F & F_FLOW) && (u & 3) == 1) { zpre = ld_prog(zcons); ypre = ld_prog(ycons); }
            if ((F & F_FLOW) && (u & 3) == 3) { zseen = max(zseen, zpre); yseen = max(yseen, ypre); }
        }
        slot = slot_n;
    }
    long long t1 = clock64();
    out[8 + lane] = xprev + a_cx + a_dv + a_bb + spins;
    if (lane == 0) { cyc[0] = t1 - t0; *(volatile int*)&ctl[5] = 1; }
}

template <int F>
int run(const char* name, double* out, long long* cyc, int bg = 0, const double* gsrc = nullptr, double* gdst = nullptr) {
    const int nblocks = 512;
    const int smem = 3 * 8192 + 32 * 65 * 8 + 16 * 512 + 64 * 16 + 16 * 512 + 1024 + 16384 + 64;
    CK(cudaFuncSetAttribute(k_step<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long hc = 0;
    for (int rep = 0; rep < 2; ++rep) {
        k_step<F><<<1, 128, smem>>>(out, cyc, nblocks, 1 << 30, bg, gsrc, gdst);
        CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
    }
    printf("%-60s bg %2d %6.1f cyc/step\n", name, bg, (double)hc / (nblocks * 8));
    return 0;
}

// ---- the same step in a ROLLED loop (runtime addressing, compact code) to see what instruction fetch costs a lone warp
template <int UNROLL>
__global__ void __launch_bounds__(128, 1) k_rolled(double* out, long long* cyc, int nsteps, int spin_iters) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    unsigned char* stage = smem;
    unsigned char* win = smem + 3 * 8192;
    unsigned char* zr = win + 32 * 65 * 8;
    unsigned char* yr = zr + 16 * 512;
    unsigned char* zo = yr + 64 * 16;
    int* ctl = (int*)(zo + 16 * 512);
    for (int i = threadIdx.x; i < 3 * 1024; i += blockDim.x) ((double*)stage)[i] = 1e-3 * (i % 7);
    for (int i = threadIdx.x; i < 32 * 65; i += blockDim.x) ((double*)win)[i] = 1.0;
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) ((ulonglong2*)zr)[i] = make_ulonglong2(0ull, (unsigned long long)INT_MAX);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) ((ulonglong2*)yr)[i] = make_ulonglong2(0ull, (unsigned long long)INT_MAX);
    if (threadIdx.x < 8) ctl[threadIdx.x] = INT_MAX;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const uint32_t stage_s = smem_u32(stage) + lane * 16, xw_lane = smem_u32(win) + lane * 65 * 8;
    const uint32_t zin_s = smem_u32(zr) + lane * 16, yin_s = smem_u32(yr);
    const uint32_t zo_s = smem_u32(zo) + lane * 16;
    const bool ypred = out[5] == 0.0 && lane == 31;
    unsigned spins = 0;
    int* status = (int*)&ctl[4];
    double xy = 0.0, acc1 = 1.0, ycur = 0.0;
    double cz = 1e-3, cy = 2e-3, cx = 3e-3, dv = 0.5, bb = 1.0;
    unsigned long long vz = 0ull, tz = INT_MAX, vy = 0ull, ty = INT_MAX;
    bool ok = true;
    // running addresses
    uint32_t sa = stage_s, wa = xw_lane, za = zin_s, ya = yin_s, oa = zo_s;
    long long t0 = clock64();
#pragma unroll UNROLL
    for (int s = 0; s < nsteps; ++s) {
        const int tag = s + 1;
        if (!ok) {
            do {
                if ((++spins & 0xfffu) == 0) { if (spins >= (unsigned)spin_iters || *(volatile int*)status != 0) return; }
                ld_s128(za, vz, tz);
                ld_s128(ya, vy, ty);
            } while (!__all_sync(FULL, (int)(unsigned)tz >= tag && (int)(unsigned)ty >= tag + 32));
        }
        // advance the running addresses (wrap with masks)
        za = zin_s + ((za - zin_s + 512) & 8191);
        ya = yin_s + ((ya - yin_s + 16) & 1023);
        unsigned long long vzn, tzn, vyn, tyn;
        ld_s128(za, vzn, tzn);
        ld_s128(ya, vyn, tyn);
        if (lane == 0) xy = ycur;
        const double acc2 = fma(-cy, xy, acc1);
        const double xnew = fma(-cz, __longlong_as_double((long long)vz), acc2);
        st_s128(oa, xnew, (7ull << 32) | (unsigned)tag);
        xy = __shfl_up_sync(FULL, xnew, 1);
        acc1 = fma(-cx, xnew, bb * dv);
        ycur = __longlong_as_double((long long)vy);
        sts_f64(wa, xnew);
        if (ypred) st_s128(oa, xnew, (7ull << 32) | (unsigned)tag);
        oa = zo_s + ((oa - zo_s + 512) & 8191);
        sa = stage_s + ((sa - stage_s + 1024) & 8191);
        wa = xw_lane + ((wa - xw_lane + 8) & 511);
        lds_2f64(sa, cz, cy);
        lds_2f64(sa + 512, cx, dv);
        bb = lds_f64(wa);
        vz = vzn, tz = tzn, vy = vyn, ty = tyn;
        ok = __all_sync(FULL, (int)(unsigned)tz >= tag + 1 && (int)(unsigned)ty >= tag + 33);
    }
    long long t1 = clock64();
    out[8 + lane] = xy + cx + dv + bb + spins;
    if (lane == 0) cyc[0] = t1 - t0;
}
template <int UNROLL>
int run_rolled(double* out, long long* cyc) {
    const int nsteps = 4096;
    const int smem = 3 * 8192 + 32 * 65 * 8 + 16 * 512 + 64 * 16 + 16 * 512 + 64;
    CK(cudaFuncSetAttribute(k_rolled<UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long hc = 0;
    for (int rep = 0; rep < 2; ++rep) {
        k_rolled<UNROLL><<<1, 128, smem>>>(out, cyc, nsteps, 1 << 30);
        CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_rolled<UNROLL>);
    printf("rolled step loop, unroll %2d: %6.1f cyc/step\n", UNROLL, (double)hc / nsteps);
    return 0;
}

int main() {
    double* out; long long* cyc;
    CK(cudaMalloc(&out, 8 * 4096)); CK(cudaMalloc(&cyc, 8 * 64));
    CK(cudaMemset(out, 0, 8 * 4096));
    constexpr int ALL = F_ZPOLL | F_YPOLL | F_ZOUT | F_YOUT | F_WIN | F_COEF | F_FLOW | F_SLOW | F_VOTE;
    run_rolled<1>(out, cyc); run_rolled<2>(out, cyc); run_rolled<4>(out, cyc); run_rolled<8>(out, cyc); run_rolled<16>(out, cyc);
    double *gsrc, *gdst;
    CK(cudaMalloc(&gsrc, 8 << 20)); CK(cudaMalloc(&gdst, 8 << 20)); CK(cudaMemset(gsrc, 0, 8 << 20));
    run<ALL>("all features (as in the kernel)", out, cyc);
    for (int bg : {1, 2, 4, 8, 15}) run<ALL>("all features + background", out, cyc, bg, gsrc, gdst);
    run<ALL | F_ZSHARED>("z-out through st.shared instead of generic", out, cyc);
    run<ALL & ~F_SLOW>("no slow path", out, cyc);
    run<(ALL & ~F_SLOW) | F_NOINL>("slow path in a noinline function", out, cyc);
    run<ALL & ~(F_SLOW | F_VOTE)>("no slow path, no vote", out, cyc);
    run<ALL & ~F_FLOW>("no flow control", out, cyc);
    run<ALL & ~F_ZOUT>("no z-out", out, cyc);
    run<ALL & ~F_YOUT>("no y-out", out, cyc);
    run<ALL & ~(F_ZOUT | F_YOUT)>("no z-out, no y-out", out, cyc);
    run<ALL & ~F_WIN>("no window store", out, cyc);
    run<ALL & ~F_COEF>("no operand loads", out, cyc);
    run<ALL & ~F_YPOLL>("no y poll", out, cyc);
    run<ALL & ~(F_YPOLL | F_ZPOLL)>("no polls", out, cyc);
    run<F_COEF>("operand loads only (+shuffle, FMAs)", out, cyc);
    run<0>("shuffle + FMAs only", out, cyc);
    run<F_COEF | F_WIN | F_ZOUT>("operands, window, z-out", out, cyc);
    run<F_COEF | F_WIN | F_ZOUT | F_ZPOLL | F_VOTE>("operands, window, z-out, z poll + vote", out, cyc);
    CK(cudaDeviceSynchronize());
    return 0;
}
