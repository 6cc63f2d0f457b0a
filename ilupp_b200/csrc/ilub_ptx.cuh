// Inline-PTX helpers: scoped acquire/release accesses, mbarrier and 1-D bulk (TMA) copies.
#pragma once
#include <cstdint>

namespace ilub {

// ---- global memory, gpu scope --------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_gpu_u8(const unsigned char* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u32(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_gpu_u8(const unsigned char* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u8(unsigned char* p, unsigned v) {
    asm volatile("st.release.gpu.global.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu_u8(unsigned char* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_gpu_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu_f64(double* p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// one architecturally atomic 16-byte access (PTX .b128 scalar type): value and epoch tag travel together,
// so the global dependency path needs no fence and no second load
__device__ __forceinline__ void st_mail(ulonglong2* p, double x, unsigned long long tag) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], q;\n\t}" ::"l"(p),
                 "l"(__double_as_longlong(x)), "l"(tag)
                 : "memory");
}
__device__ __forceinline__ bool ld_mail(const ulonglong2* p, unsigned long long tag, double& x) {
    long long xb;
    unsigned long long t;
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}"
                 : "=l"(xb), "=l"(t)
                 : "l"(p)
                 : "memory");
    x = __longlong_as_double(xb);
    return t == tag;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// ---- shared memory, cta scope ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int ld_acquire_cta_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_cta_s32(const int* p) {
    int v;
    asm volatile("ld.relaxed.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_s32(int* p, int v) {
    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_cta_s32(int* p, int v) {
    asm volatile("st.relaxed.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// x ring entries {x bits, tag}: one atomic 16-byte shared-memory access, no flag array, no fence
__device__ __forceinline__ void st_ring(ulonglong2* p, double x, unsigned long long tag) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.cta.shared.b128 [%0], q;\n\t}" ::"r"(
                     smem_u32(p)),
                 "l"(__double_as_longlong(x)), "l"(tag)
                 : "memory");
}
__device__ __forceinline__ void st_ring_s(uint32_t saddr, double x, unsigned long long tag) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.cta.shared.b128 [%0], q;\n\t}" ::"r"(saddr),
                 "l"(__double_as_longlong(x)), "l"(tag)
                 : "memory");
}
__device__ __forceinline__ bool ld_ring(const ulonglong2* p, unsigned long long tag, double& x) {
    long long xb;
    unsigned long long t;
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.cta.shared.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}"
                 : "=l"(xb), "=l"(t)
                 : "r"(smem_u32(p))
                 : "memory");
    x = __longlong_as_double(xb);
    return t == tag;
}
__device__ __forceinline__ void ld_b128_cta(uint32_t saddr, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.cta.shared.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}"
                 : "=l"(lo), "=l"(hi)
                 : "r"(saddr)
                 : "memory");
}
__device__ __forceinline__ void ld_b128_gpu(const ulonglong2* p, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}"
                 : "=l"(lo), "=l"(hi)
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

// ---- mbarrier + bulk async copy (TMA engine, 1-D) ----------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// plain arrival (release at cta scope): completes the current phase of a count-1 barrier
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// global -> shared bulk copy completing on an mbarrier; size multiple of 16, both addresses 16-B aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- arithmetic in the reference's operation order ---------------------------------------------
// EXACT: separately rounded multiply and subtract/add and a true division, i.e. what the reference
// executes when compiled without FMA contraction.  !EXACT: fused multiply-add.
template <bool EXACT>
__device__ __forceinline__ double sub_mul(double acc, double a, double b) {  // acc - a*b
    if (EXACT) return __dsub_rn(acc, __dmul_rn(a, b));
    return fma(-a, b, acc);
}
template <bool EXACT>
__device__ __forceinline__ double add_mul(double acc, double a, double b) {  // acc + a*b
    if (EXACT) return __dadd_rn(acc, __dmul_rn(a, b));
    return fma(a, b, acc);
}

}  // namespace ilub
