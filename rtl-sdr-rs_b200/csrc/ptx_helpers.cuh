// ptx_helpers.cuh — device-side PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP).
// Compiles under nvcc and under NVRTC (csrc/rtc.cpp hands this file to the runtime compiler as an in-memory header).
#pragma once
#ifdef __CUDACC_RTC__
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef long long int64_t;
#else
#include <cstdint>
#endif

namespace sdr {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    // make the init visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; src/dst 16-B aligned, bytes % 16 == 0; completes on `bar`.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 16-byte asynchronous copy global -> shared by one thread (SASS LDGSTS), bypassing L1
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc)
                 : "memory");
}
// the mbarrier receives one arrival from this thread when all of its cp.async issued so far have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// same, with an L2 evict-first policy: the IQ stream is read exactly once
__device__ __forceinline__ void bulk_g2s_stream(void *smem_dst, const void *gsrc, uint32_t bytes,
                                                uint64_t *bar) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}

// atan2 for the f32 discriminators: minimax odd polynomial of degree 15 on [0, 1] (max error 1.2e-7 rad in f32) after the
// usual octant reduction; ~20 instructions instead of atan2f's ~40.  Not both arguments zero.
__device__ __forceinline__ float poly_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float q = __fdividef(mn, mx);
    const float s = q * q;
    float p = -0.004054520279169083f;
    p = fmaf(p, s, 0.021862786263227463f);
    p = fmaf(p, s, -0.0559120811522007f);
    p = fmaf(p, s, 0.09642180055379868f);
    p = fmaf(p, s, -0.13908623158931732f);
    p = fmaf(p, s, 0.19946564733982086f);
    p = fmaf(p, s, -0.33329859375953674f);
    p = fmaf(p, s, 0.9999993443489075f);
    p *= q;
    if (ay > ax) p = 1.57079632679489662f - p;
    if (x < 0.f) p = 3.14159265358979324f - p;
    return copysignf(p, y);
}

}  // namespace sdr
