// Fused bidirectional cost-volume half (CrossLayerLight.cross, pointconv_util.py:1826-1850) for sm_100a.
//
// Reference op chain per call: knn_point -> 2x index_points_group (grouping_operation + permutes) ->
// subtract -> Conv2d(3->D) -> `repeat` copy of points1 -> 2 adds -> ReLU/LeakyReLU -> Conv2d(D->D')
// -> LeakyReLU -> F.max_pool2d over the K neighbours: six passes over [B,D,K,N] tensors (268 MB each
// at the 8192-point level).  Here NONE of the [B,N,K,*] tensors reaches HBM:
//   * producers (two groups of 4 warps, one thread per (point, neighbour) row; 128-row tile = 4 points x
//     K=32 neighbours) gather points2[idx], add points1 and the positional encoding
//     pos_w (xyz2[idx]-xyz1) + pos_b, apply the activation, split fp32 -> bf16 hi/lo and write the
//     A-operand tile in the swizzled layout;
//   * Conv2d(D->D') runs on tcgen05 (3 MMAs per K-step, fp32 accumulation in TMEM);
//   * epilogue: one warp owns the 32 TMEM lanes (= the 32 neighbours) of ONE point, so the max over K
//     is a warp-wide redux.sync per output channel; bias + LeakyReLU commute with the max (monotone) and
//     are applied to the single surviving value; only [B,N,D'] is written.
#include "tc_gemm.cuh"

namespace kdpc {
namespace tc {

constexpr int CV_K = 32;            // neighbours per point: one epilogue warp per point
constexpr int CV_MAX_D = 256;

struct CostVolProducer {
    static constexpr int kWarps = 8, kGroups = 2;
    struct Args {
        const float *xyz1;   // [B,S,3] queries
        const float *xyz2;   // [B,N,3] candidates
        const float *p1;     // [B,S,D]
        const float *p2;     // [B,N,D]
        const int *idx;      // [B,S,32]
        const float *pos_w;  // [D,3]
        const float *pos_b;  // [D]
        int s, n, d;
        float slope;         // activation after the sum (0 = ReLU)
    };
    // (w0,w1,w2,b) per channel, staged once per CTA
    static __device__ __forceinline__ float4 *posw() {
        __shared__ float4 posw_s[CV_MAX_D];
        return posw_s;
    }
    static __device__ __forceinline__ void prologue(const Args &a, int tid, int nthreads) {
        float4 *pw = posw();
        for (int c = tid; c < a.d; c += nthreads)
            pw[c] = make_float4(__ldg(a.pos_w + c * 3), __ldg(a.pos_w + c * 3 + 1), __ldg(a.pos_w + c * 3 + 2), __ldg(a.pos_b + c));
    }
    const Args &a;
    const GemmShape &g;
    const float *p2row, *p1row;
    float dx, dy, dz;

    __device__ CostVolProducer(const Args &a_, const GemmShape &g_) : a(a_), g(g_) {}

    __device__ __forceinline__ void begin_tile(long long tile, int r) {
        long long row = tile * TILE_M + r;                   // row = (b*S + i)*32 + k
        if (row >= g.m) row = g.m - 1;                       // padded rows repeat the last row; never stored
        const long long pt = row >> 5;
        const long long b = pt / a.s;
        const int j = __ldg(a.idx + row);
        const float *q = a.xyz1 + pt * 3;
        const float *c = a.xyz2 + (b * a.n + j) * 3;
        dx = __ldg(c) - __ldg(q);
        dy = __ldg(c + 1) - __ldg(q + 1);
        dz = __ldg(c + 2) - __ldg(q + 2);
        p2row = a.p2 + (b * a.n + j) * (long long)a.d;
        p1row = a.p1 + pt * (long long)a.d;
    }

    __device__ __forceinline__ void fill(int chunk, unsigned char *a_hi, unsigned char *a_lo, int r) {
        const float4 *pw = posw();
        const int c0 = chunk * CHUNK_K;
        const int units = min(8, (a.d - c0) >> 3);           // d % 8 == 0 (checked on the host)
        float4 g2[16];                                       // the gathered row chunk: all loads in flight at once
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (u < units) {
                g2[2 * u] = __ldg(reinterpret_cast<const float4 *>(p2row + c0 + u * 8));
                g2[2 * u + 1] = __ldg(reinterpret_cast<const float4 *>(p2row + c0 + u * 8 + 4));
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
            if (u < units) {
                const float va[8] = {g2[2 * u].x, g2[2 * u].y, g2[2 * u].z, g2[2 * u].w,
                                     g2[2 * u + 1].x, g2[2 * u + 1].y, g2[2 * u + 1].z, g2[2 * u + 1].w};
                // points1 row of this point: same address for the 32 lanes of the warp (L1 broadcast)
                const float4 q0 = __ldg(reinterpret_cast<const float4 *>(p1row + c0 + u * 8));
                const float4 q1 = __ldg(reinterpret_cast<const float4 *>(p1row + c0 + u * 8 + 4));
                const float vb[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float4 w = pw[c0 + u * 8 + e];
                    const float t = (va[e] + vb[e]) + (w.w + w.x * dx + w.y * dy + w.z * dz);   // pointconv_util.py:1843
                    v[e] = t > 0.f ? t : t * a.slope;
                }
                split8(v, hi, lo);
            }
            const uint32_t off = sw128_offset(r, u);
            *reinterpret_cast<uint4 *>(a_hi + off) = hi;
            *reinterpret_cast<uint4 *>(a_lo + off) = lo;
        }
    }
};

// order-preserving float <-> signed int (so that redux.sync.max.s32 is a float max)
__device__ __forceinline__ int f2ord(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

struct MaxKEpilogue {
    struct Args {
        const float *bias;    // [n] or nullptr
        float slope;          // LeakyReLU after the conv
        float *out;           // [points, ldo]
        int ldo;
        long long points;
    };
    __device__ __forceinline__ void tile(const Args &e, const GemmShape &g, long long tile, uint32_t t_acc, int quarter,
                                         int lane) const {
        const long long pt = tile * (TILE_M / CV_K) + quarter;
        for (int c0 = 0; c0 < g.n_pad; c0 += 32) {
            float v[32];
            tmem_ld_32x32(t_acc + (uint32_t)c0, v);               // lane = neighbour k, v[j] = channel c0+j
            int mine = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int m = __reduce_max_sync(0xffffffffu, f2ord(v[j]));
                if (lane == j) mine = m;
            }
            const int col = c0 + lane;
            if (pt < e.points && col < g.n) {
                float y = ord2f(mine);
                if (e.bias) y += __ldg(e.bias + col);
                y = y > 0.f ? y : y * e.slope;
                e.out[pt * e.ldo + col] = y;
            }
        }
    }
};

}  // namespace tc
}  // namespace kdpc

using namespace kdpc;
using namespace kdpc::tc;

KDPC_API int kdpc_costvol_fused(int b, int s, int n, int k, int d, int d_out, const float *xyz1, const float *xyz2,
                                const float *p1, const float *p2, const int *idx, const float *pos_w,
                                const float *pos_b, float slope_pre, const void *wpacked, const float *bias,
                                float slope_post, float *out, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(xyz1 && xyz2 && p1 && p2 && idx && pos_w && pos_b && wpacked && out && b > 0 && s > 0 && n > 0 &&
                    d > 0 && d_out > 0);
    if (k != CV_K || d > CV_MAX_D || (d & 7) != 0 || d_out > 256) return KDPC_EUNSUPPORTED;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2) | reinterpret_cast<uintptr_t>(wpacked);
    if (al % 16 != 0) return KDPC_EINVAL;
    const long long points = (long long)b * s;
    GemmShape g = make_shape(points * CV_K, d_out, d, wpacked);
    const size_t smem = smem_bytes(g.n_pad, g.stages);
    auto kern = tc_gemm_kernel<CostVolProducer, MaxKEpilogue>;
    KDPC_ENSURE_SMEM(kern, 201 * 1024);
    CostVolProducer::Args pa{xyz1, xyz2, p1, p2, idx, pos_w, pos_b, s, n, d, slope_pre};
    MaxKEpilogue::Args ea{bias, slope_post, out, d_out, points};
    const unsigned grid = (unsigned)(g.num_tiles < kNumSMs ? g.num_tiles : kNumSMs);
    kern<<<grid, num_threads<CostVolProducer>(), smem, to_stream(stream)>>>(g, pa, ea);
    KDPC_RETURN_LAST();
}
