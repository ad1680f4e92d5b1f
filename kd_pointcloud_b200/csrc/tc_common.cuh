// tcgen05 / TMEM helpers shared by the tensor-core kernels (sm_100a only).
//
// Operand convention used by every kernel here:
//   * A (activations) and B (weights) tiles are K-major bf16 in shared memory in the canonical
//     SWIZZLE_128B layout: one row = 64 bf16 = 128 bytes, the eight 16-byte units of a row are
//     XOR-swizzled with (row & 7), 8-row groups are 1024 bytes apart.  Tile bases are 1024-byte
//     aligned.  One such tile covers a K-chunk of 64.
//   * fp32 accuracy from bf16 tensor cores: every fp32 value v is split as hi = bf16(v),
//     lo = bf16(v - hi) and the product is accumulated as hi*hi + hi*lo + lo*hi in the fp32 TMEM
//     accumulator (the dropped lo*lo term is ~2^-18 relative): three MMAs per K-step, measured
//     against an fp64 reference at < 1e-5 relative (tests/test_tc_gpu.py).
//   * D lives in TMEM: row r of the 128-row tile is TMEM lane r, column n is TMEM column base+n.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace kdpc {
namespace tc {

constexpr int TILE_M = 128;          // rows per CTA tile (UMMA M, cta_group::1)
constexpr int CHUNK_K = 64;          // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;           // K per tcgen05.mma for 16-bit inputs
constexpr int A_PART_BYTES = TILE_M * 128;               // one 128x64 bf16 tile (hi or lo)
constexpr int A_STAGE_BYTES = 2 * A_PART_BYTES;          // hi + lo

// ---- descriptors -------------------------------------------------------------------------
// Instruction descriptor, kind::f16 (cute/arch/mma_sm100_desc.hpp InstrDescriptor):
//   [4,6) c_format = 1 (F32) | [7,10) a_format = 1 (BF16) | [10,13) b_format = 1 (BF16)
//   [15] a_major = 0 (K) | [16] b_major = 0 (K) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// Shared-memory matrix descriptor, K-major SWIZZLE_128B (SmemDescriptor):
//   [0,14) addr >> 4 | [16,30) LBO >> 4 = 1 (unused for swizzled K-major) | [32,46) SBO >> 4 = 64 (1024 B)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// ---- tcgen05 wrappers -----------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of the (converged) warp: the predicate ptxas recognises for single-thread tcgen05 issue
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrive when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp gets lane (base_lane + i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16-byte asynchronous global -> shared copy (LDGSTS, L2 only) and its completion hook on an mbarrier
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- fp32 -> (bf16 hi, bf16 lo) split, 8 values -> one 16-byte unit each ------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {          // a -> low half, b -> high half
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ void split8(const float (&v)[8], uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        h[i] = pack_bf16x2(a, b);
        const float ah = __uint_as_float(h[i] << 16), bh = __uint_as_float(h[i] & 0xffff0000u);
        l[i] = pack_bf16x2(a - ah, b - bh);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// byte offset of 16-byte unit `u` (0..7) of row `r` inside a SWIZZLE_128B K-major tile
__device__ __forceinline__ uint32_t sw128_offset(int r, int u) { return (uint32_t)r * 128u + (uint32_t)((u ^ (r & 7)) << 4); }

}  // namespace tc
}  // namespace kdpc
