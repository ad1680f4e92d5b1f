// Fused PointConv for sm_100a: neighbour gather + relative xyz + WeightNet + sum over K + Linear
// (+ eval BatchNorm + LeakyReLU) in ONE kernel.
//
// Reference op chain (PointConv.forward, pointconv_util.py:231-258; SceneFlowEstimatorResidual
// calls it twice per level with K = 9, :2217-2225): kNN -> 2x grouping_operation -> subtract -> cat
// ([B,N,K,3+D], 309 MB at flow0) -> three cuDNN 1x1 convs (WeightNet) -> bmm of B*N tiny matrices
// ([B,N,16(3+D)], 550 MB at flow0) -> nn.Linear (fp32 SIMT GEMM) -> BatchNorm1d -> LeakyReLU.
// Here neither the grouped tensor nor the [B,N,16C] tensor ever reaches HBM:
//   * 8 producer warps: thread (row, half) keeps its point's WeightNet outputs wn[k][8] in REGISTERS
//     for the whole tile (WeightNet weights come from the kernel-parameter constant bank, so every
//     FFMA takes its weight operand directly), gathers the neighbours' feature rows as aligned
//     float4 (4 channels = one 64-wide K-chunk: 4 channels x 16 weightnet outputs), accumulates
//     agg[ch][w] += g[k][ch] * wn[k][w] and writes the bf16 hi/lo split straight into the swizzled
//     A-operand tile in shared memory;
//   * the Linear runs on tcgen05 (tc_gemm.cuh) with the weight pre-permuted to the producer's
//     channel order [features, dx, dy, dz, 0] (linear_tc.cu: pack mode 1);
//   * bias / BatchNorm(eval) / LeakyReLU in the TMEM epilogue.
#include <cstring>
#include "tc_gemm.cuh"

namespace kdpc {
namespace tc {

// KN neighbours in NPASS passes of NB = KN / NPASS: the aggregation is linear in the neighbours, so pass p
// contributes sum_{k in pass p} g_k (x) wn_k as its own run of K-chunks accumulated into the same TMEM tile (the
// weight chunks repeat).  K = 16 (PointConvD) runs as 2 x 8, which keeps wn[NB][8] in registers.
template <int KN, int NPASS>
struct PointConvProducer {
    static constexpr int kWarps = 8, kGroups = 1;
    static constexpr bool kAsync = false;
    static constexpr int kIssuers = 0, kLookahead = 0;
    static constexpr int NB = KN / NPASS;
    struct Args {
        const float *cand_xyz;    // [B,N,3]
        const float *query_xyz;   // [B,S,3]
        const float *feats;       // [B,N,D]
        const int *idx;           // [B,S,KN]
        int n_cand, s, d;
        float w1[24], b1[8], w2[64], b2[8], w3[128], b3[16];    // WeightNet 3 -> 8 -> 8 -> 16, ReLU after each
    };
    static __device__ __forceinline__ void prologue(const Args &, int, int) {}
    const Args &a;
    const GemmShape &g;
    float wn[NB][8];
    int nb[NB];
    const float *fbase, *cbase;
    const int *ip;
    float qx, qy, qz;
    int half, cur_pass;

    __device__ PointConvProducer(const Args &a_, const GemmShape &g_) : a(a_), g(g_) {}

    __device__ __forceinline__ void begin_tile(long long tile, int ptid) {
        const int r = ptid & 127;
        half = ptid >> 7;
        long long row = tile * TILE_M + r;
        if (row >= g.m) row = g.m - 1;                       // padded rows recompute the last point; never stored
        const long long b = row / a.s;
        ip = a.idx + row * KN;
        const float *qp = a.query_xyz + row * 3;
        qx = qp[0]; qy = qp[1]; qz = qp[2];
        cbase = a.cand_xyz + b * a.n_cand * 3;
        fbase = a.feats + b * (long long)a.n_cand * a.d;
        cur_pass = -1;
    }

    // neighbour indices + this thread's 8 WeightNet outputs for the neighbours of one pass
    __device__ __forceinline__ void load_pass(int pass) {
        cur_pass = pass;
#pragma unroll
        for (int k = 0; k < NB; ++k) nb[k] = __ldg(ip + pass * NB + k);
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const float *cp = cbase + (long long)nb[k] * 3;
            const float dx = __ldg(cp) - qx, dy = __ldg(cp + 1) - qy, dz = __ldg(cp + 2) - qz;
            float h1[8], h2[8];
#pragma unroll
            for (int o = 0; o < 8; ++o)
                h1[o] = fmaxf(a.b1[o] + a.w1[o * 3 + 0] * dx + a.w1[o * 3 + 1] * dy + a.w1[o * 3 + 2] * dz, 0.f);
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                float t = a.b2[o];
#pragma unroll
                for (int i = 0; i < 8; ++i) t += a.w2[o * 8 + i] * h1[i];
                h2[o] = fmaxf(t, 0.f);
            }
            if (half == 0) {
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    float t = a.b3[o];
#pragma unroll
                    for (int i = 0; i < 8; ++i) t += a.w3[o * 8 + i] * h2[i];
                    wn[k][o] = fmaxf(t, 0.f);
                }
            } else {
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    float t = a.b3[8 + o];
#pragma unroll
                    for (int i = 0; i < 8; ++i) t += a.w3[(8 + o) * 8 + i] * h2[i];
                    wn[k][o] = fmaxf(t, 0.f);
                }
            }
        }
    }

    __device__ __forceinline__ void fill(int chunk, unsigned char *a_hi, unsigned char *a_lo, int ptid) {
        const int r = ptid & 127;
        const int pass = chunk / g.wchunks;
        chunk -= pass * g.wchunks;
        if (pass != cur_pass) load_pass(pass);
        float acc[4][8];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
        if (chunk * 4 < a.d) {
            const float *fp = fbase + chunk * 4;
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(fp + (long long)nb[k] * a.d));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[0][j] = fmaf(v.x, wn[k][j], acc[0][j]);
                    acc[1][j] = fmaf(v.y, wn[k][j], acc[1][j]);
                    acc[2][j] = fmaf(v.z, wn[k][j], acc[2][j]);
                    acc[3][j] = fmaf(v.w, wn[k][j], acc[3][j]);
                }
            }
        } else {                                             // last chunk: channels (dx, dy, dz, 0)
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const float *cp = cbase + (long long)nb[k] * 3;
                const float dx = __ldg(cp) - qx, dy = __ldg(cp + 1) - qy, dz = __ldg(cp + 2) - qz;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[0][j] = fmaf(dx, wn[k][j], acc[0][j]);
                    acc[1][j] = fmaf(dy, wn[k][j], acc[1][j]);
                    acc[2][j] = fmaf(dz, wn[k][j], acc[2][j]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 hi, lo;
            split8(acc[c], hi, lo);
            const uint32_t off = sw128_offset(r, c * 2 + half);
            *reinterpret_cast<uint4 *>(a_hi + off) = hi;
            *reinterpret_cast<uint4 *>(a_lo + off) = lo;
        }
    }
};

// (An asynchronous cp.async variant of the gathers was measured 30 % SLOWER: it bypasses L1, and eight consecutive
// K-chunks share each 128-byte line of a neighbour row - the synchronous __ldg path lives on those L1 hits.)
static int kdpc_pointconv_stages = 2;

template <int KN, int NPASS>
static int launch_pointconv(long long m, int n_out, const typename PointConvProducer<KN, NPASS>::Args &pa, const void *wpacked,
                            StoreEpilogue::Args ea, void *ws, cudaStream_t st) {
    using P = PointConvProducer<KN, NPASS>;
    GemmShape g = make_shape(m, n_out, (pa.d + 4) * 16, wpacked);
    g.num_chunks = g.wchunks * NPASS;                        // one run of the weight's K-chunks per neighbour pass
    g.chunks_per_split = g.num_chunks;
    if (ws != nullptr) plan_split_k(g);
    ea.partial = reinterpret_cast<float *>(ws);
    // the neighbour gathers live on L1 hits (8 consecutive K-chunks share a 128-byte line): two operand stages are
    // enough to keep the MMA fed and leave ~100 KB of the unified L1/shared array to the cache
    if (g.stages > kdpc_pointconv_stages) g.stages = kdpc_pointconv_stages;
    const size_t smem = smem_bytes(g.n_pad, g.stages);
    auto kern = tc_gemm_kernel<P, StoreEpilogue>;
    KDPC_ENSURE_SMEM(kern, 201 * 1024);
    const long long work = g.num_tiles * g.splits;
    const unsigned grid = (unsigned)(work < kNumSMs ? work : kNumSMs);
    kern<<<grid, num_threads<P>(), smem, st>>>(g, pa, ea);
    if (g.splits > 1) return launch_splitk_reduce(g, ea, st);
    return (int)cudaGetLastError();
}

static GemmShape pointconv_shape(long long m, int n_out, int d, int k) {
    GemmShape g = make_shape(m, n_out, (d + 4) * 16, nullptr);
    g.num_chunks = g.wchunks * (k == 16 ? 2 : 1);
    g.chunks_per_split = g.num_chunks;
    plan_split_k(g);
    return g;
}

}  // namespace tc
}  // namespace kdpc

using namespace kdpc;
using namespace kdpc::tc;

KDPC_API void kdpc_pointconv_set_stages(int n) { kdpc::tc::kdpc_pointconv_stages = n < 2 ? 2 : n; }

KDPC_API long long kdpc_pointconv_fused_ws_bytes(int b, int s, int k, int d, int n_out) {
    if (b <= 0 || s <= 0 || d <= 0 || n_out <= 0 || n_out > 256) return 0;
    return (long long)split_k_ws_bytes(pointconv_shape((long long)b * s, n_out, d, k));
}

KDPC_API int kdpc_pointconv_fused(int b, int n, int s, int k, int d, int n_out, const float *cand_xyz,
                                  const float *query_xyz, const float *feats, const int *idx,
                                  const float *wn_params /* host: w1[24] b1[8] w2[64] b2[8] w3[128] b3[16] */,
                                  const void *wpacked, const float *scale, const float *shift, float slope,
                                  void *ws, float *out, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(cand_xyz && query_xyz && feats && idx && wn_params && wpacked && out && b > 0 && n > 0 && s > 0 &&
                    d > 0 && n_out > 0);
    if (n_out > 256 || (d & 3) != 0 || (k != 9 && k != 16)) return KDPC_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(feats) % 16) != 0 || (reinterpret_cast<uintptr_t>(out) % 16) != 0 ||
        (reinterpret_cast<uintptr_t>(wpacked) % 16) != 0 || (reinterpret_cast<uintptr_t>(ws) % 16) != 0)
        return KDPC_EINVAL;
    PointConvProducer<9, 1>::Args pa;                        // (the Args layout does not depend on KN / NPASS)
    pa.cand_xyz = cand_xyz; pa.query_xyz = query_xyz; pa.feats = feats; pa.idx = idx;
    pa.n_cand = n; pa.s = s; pa.d = d;
    const float *p = wn_params;
    for (int i = 0; i < 24; ++i) pa.w1[i] = *p++;
    for (int i = 0; i < 8; ++i) pa.b1[i] = *p++;
    for (int i = 0; i < 64; ++i) pa.w2[i] = *p++;
    for (int i = 0; i < 8; ++i) pa.b2[i] = *p++;
    for (int i = 0; i < 128; ++i) pa.w3[i] = *p++;
    for (int i = 0; i < 16; ++i) pa.b3[i] = *p++;
    StoreEpilogue::Args ea{scale, shift, slope, 1.f, 0.f, nullptr, out, n_out, nullptr};
    if (k == 9) return launch_pointconv<9, 1>((long long)b * s, n_out, pa, wpacked, ea, ws, to_stream(stream));
    PointConvProducer<16, 2>::Args pb;
    static_assert(sizeof(pb) == sizeof(pa), "Args layout");
    memcpy(&pb, &pa, sizeof(pa));
    return launch_pointconv<16, 2>((long long)b * s, n_out, pb, wpacked, ea, ws, to_stream(stream));
}
