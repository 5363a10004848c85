// Shared device helpers for the DCL-Net B200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "dcl_b200 kernels are written for sm_100a only"
#endif

#define DCL_API extern "C" __attribute__((visibility("default")))
#define DCL_DIVUP(a, b) (((a) + (b) - 1) / (b))

#define DCL_RETURN_IF_BAD(cond)                       \
    do {                                              \
        if (!(cond)) return (int)cudaErrorInvalidValue; \
    } while (0)

// Diagnostic only: number of kernels this library has launched in the process
// (dcl_b200_launch_count()).  Not used by any computation.
extern unsigned long long g_dcl_kernel_launches;
#define DCL_COUNT_LAUNCHES(n) (g_dcl_kernel_launches += (unsigned long long)(n))

// Called after an entry point has issued `n_kernels` launches.
static inline int dcl_launch_status(int n_kernels = 1) {
    DCL_COUNT_LAUNCHES(n_kernels);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Reference arithmetic.  nvcc (-O2, default -fmad=true) contracts the
// reference's  dx*dx + dy*dy + dz*dz  into  fma(dz,dz, fma(dx,dx, dy*dy))
// (checked in the SASS of oracle/_ref).  We spell the
// order out so that the result never depends on compiler contraction choices.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float dcl_dist2(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx);
    const float dy = __fsub_rn(ay, by);
    const float dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// w0*f0 + w1*f1 + w2*f2 as the reference's SASS evaluates it:
// fma(w2,f2, fma(w0,f0, w1*f1))  (SURVEY.md Appendix A.0).
__device__ __forceinline__ float dcl_interp3(float w0, float f0, float w1, float f1, float w2, float f2) {
    return __fmaf_rn(w2, f2, __fmaf_rn(w0, f0, __fmul_rn(w1, f1)));
}

// ---------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dcl_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void dcl_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dcl_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void dcl_fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void dcl_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void dcl_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dcl_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void dcl_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dcl_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool dcl_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(dcl_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void dcl_mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!dcl_mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-B aligned.
__device__ __forceinline__ void dcl_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            dcl_smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(dcl_smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------
// Streaming (read-once / write-once) global accesses
// ---------------------------------------------------------------------------
__device__ __forceinline__ void dcl_st_stream_f4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void dcl_st_stream_f1(float* p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ int4 dcl_ld_stream_i4(const int* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float4 dcl_ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
