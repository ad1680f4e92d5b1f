// Shared device helpers for libkdpc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/kdpc.h"

#define KDPC_API extern "C" __attribute__((visibility("default")))

#define KDPC_CHECK_ARGS(cond) do { if (!(cond)) return KDPC_EINVAL; } while (0)
#define KDPC_RETURN_LAST() return (int)cudaGetLastError()

// Opt in to > 48 KB dynamic shared memory ONCE per (kernel, device): keeps cudaFuncSetAttribute out of
// the steady state (and out of CUDA-graph capture after the warm-up run).
#define KDPC_ENSURE_SMEM(kern, bytes)                                                               \
    do {                                                                                            \
        static int done_[64];                                                                       \
        int dev_ = 0;                                                                               \
        cudaGetDevice(&dev_);                                                                       \
        if (dev_ < 0 || dev_ >= 64 || !done_[dev_]) {                                               \
            cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
            if (e_ != cudaSuccess) return (int)e_;                                                  \
            if (dev_ >= 0 && dev_ < 64) done_[dev_] = 1;                                            \
        }                                                                                           \
    } while (0)

static inline cudaStream_t to_stream(kdpc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline long long div_up_ll(long long a, long long b) { return (a + b - 1) / b; }

namespace kdpc {

// SM count of the CURRENT device, queried once per device (148 on a B200): grids of the persistent kernels and the
// split-K / queries-per-warp plans are sized from it.
// kdpc_set_sm_limit(n) caps it: persistent kernels then launch at most n CTAs, leaving the other SMs to kernels of a
// concurrent stream (runner.FlowRunner pipelines the sampling pyramid of the next batch beside the current one).
// Work PLANS (split-K, splits over rows, queries per warp) always use the real count, so results do not depend on the cap.
extern "C" int kdpc_sm_limit(void);
static inline int device_sms() {
    static int cached[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}
static inline int num_sms() {                    // grid size of the persistent kernels
    const int n = device_sms(), lim = kdpc_sm_limit();
    return (lim > 0 && lim < n) ? lim : n;
}

// ---- exact-rounding fp32 helpers: never contracted or re-associated by nvcc ----------------
__device__ __forceinline__ float sq_norm3(float x, float y, float z) {
    // torch: sum(p ** 2, -1)  ->  rn(rn(rn(x*x) + rn(y*y)) + rn(z*z))      (pointconv_util.py:92-93)
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
__device__ __forceinline__ float expansion_dist(float qx, float qy, float qz, float qq,
                                                float cx, float cy, float cz, float cc) {
    // torch: -2 * matmul(src, dst^T) + |src|^2 + |dst|^2                     (pointconv_util.py:91-93)
    float dot = __fmaf_rn(qz, cz, __fmaf_rn(qy, cy, __fmul_rn(qx, cx)));
    return __fadd_rn(__fmaf_rn(-2.f, dot, qq), cc);
}
__device__ __forceinline__ float direct_dist(float dx, float dy, float dz) {
    // reference CUDA kernels after nvcc contraction: fma(dz,dz, fma(dx,dx, rn(dy*dy)))
    // (sampling_gpu.cu:133, interpolate_gpu.cu:37, ball_query_gpu.cu:33)
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    // bounded: a lost TMA completion traps (launch error) instead of hanging the GPU
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// streaming 128-bit accesses (data touched once: keep it out of L1)
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace kdpc
