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
//     channel order [dx, dy, dz, 0, features] (linear_tc.cu: pack mode 1);
//   * bias / BatchNorm(eval) / LeakyReLU in the TMEM epilogue.
#include <cstring>
#include "tc_gemm.cuh"

namespace kdpc {
namespace tc {

// KN neighbours in NPASS passes of NB = KN / NPASS: the aggregation is linear in the neighbours, so pass p
// contributes sum_{k in pass p} g_k (x) wn_k as its own run of K-chunks accumulated into the same TMEM tile (the
// weight chunks repeat).  K = 16 (PointConvD) runs as 2 x 8, which keeps wn[NB][8] in registers.
// K-chunk 0 of a pass = channels (dx, dy, dz, 0); chunk i >= 1 = feature channels 4(i-1) .. 4(i-1)+3.
//
// Neighbour gathers: straight into registers, feature chunks in PAIRS - the two half-lanes of a row are adjacent lanes
// and fetch the two 16-byte halves of each neighbour's 32-byte sector once per pair (a warp request touches 16 rows,
// one L1 line access per row and pair), then hand each other the half the current chunk needs by shuffle: 36 SHFL per
// chunk buy half the L1 line accesses, the kernel's real limiter (-9 % at flow0, and the natural row order became as
// fast as the Morton order).  Four consecutive chunk pairs share each 128-byte line (L1 hits), and with a Morton row
// order the 128 queries of a tile share most of their neighbours.
// Rejected by measurement on the same box (same results, tools/bench_pointconv.py, tools/trace_pointconv.py):
//   * cp.async staging of the gathers in shared memory (per CTA, per row quarter and per warp; double buffered; also
//     interleaved into the FFMA loop): never faster - a warp-level LDGSTS whose lanes touch 16 lines blocks the
//     issuing warp for ~150 cycles, and the 74 KB of staging take the L1 away;
//   * software pipelining (chunk c+1's loads issued after chunk c's FFMAs, stage acquired after the FFMAs): 5 %
//     slower, the extra live registers spill in the WeightNet phase; refilling neighbour k's registers with the next
//     pair's gather right after their last shuffle (no extra live registers on paper): ptxas hoists the loads and
//     spills inside the loop (34 LDL per pair), 206 -> 243 us;
//   * prefetch.global.L1 of every neighbour row's next 128-byte line two to six chunks ahead: 5 % slower;
//   * 4 threads per row (16 producer warps at 96 registers, wn[9][4] each) for more latency hiding: 15 % slower - the
//     kernel is bound by L1 line accesses (every warp request touches 16 neighbour rows = 16 wavefronts, 1152 per
//     chunk) and by issue slots, not by exposed latency; the redundant hidden-layer work and 8-byte stores cost more.
template <int KN, int NPASS>
struct PointConvProducer {
    static constexpr int kWarps = 8, kGroups = 1;
    static constexpr bool kAsync = false;
    // 12 warps (the MMA issuer lives in the first epilogue warp): 168 registers per thread, so wn[NB][8] really
    // stays in registers (at 128 it was spilled and re-read from local memory every K-chunk: 85 LDL per chunk,
    // 1.1 GB of L2 traffic per launch at flow0)
    static constexpr bool kMergedIssuer = true;
    static constexpr bool kOwnsLoop = true;                  // run_tile() below instead of fill()
    static constexpr int kIssuers = 0, kLookahead = 0;
    static constexpr int NB = KN / NPASS;
    struct Args {
        const float *cand_xyz;    // [B,N,3]
        const float *query_xyz;   // [B,S,3]
        const float *feats;       // [B,N,D]
        const int *idx;           // [B,S,KN]
        int n_cand, s, d;
        const int *order;         // optional [B] x order_stride: tile position i of cloud b processes query order[b][i]
        int order_stride;
        const float *wn_pre;      // optional [B*S, KN, 16]: the WeightNet outputs, precomputed by pointconv_weightnet_kernel
        float w1[24], b1[8], w2[64], b2[8], w3[128], b3[16];    // WeightNet 3 -> 8 -> 8 -> 16, ReLU after each
    };
    static __device__ __forceinline__ void prologue(const Args &, int, int) {}
    const Args &a;
    const GemmShape &g;
    float wn[NB][8];
    uint32_t nb[NB];              // element offset of each neighbour's feature row from a.feats
    const int *ip;
    float qx, qy, qz;
    int half, boff;

    __device__ PointConvProducer(const Args &a_, const GemmShape &g_) : a(a_), g(g_) {}

    // Thread mapping: producer warp w owns tile rows 16w .. 16w+15, lane = 2 * (row & 15) + half.  A warp is then the
    // only reader of its rows' gathered data (no cross-warp hand-over of staging buffers), the two half-lanes of a row
    // read the same 16 bytes (one broadcast shared-memory word / one coalesced global request), and lane (row, h)
    // fetches piece h of its own row's neighbours - no index exchange.
    // Which of the warp's 16 rows a lane pair owns: a 16-byte shared-memory store is processed per QUARTER warp (8 lanes =
    // 4 rows x 2 halves), and in the SWIZZLE_128B layout rows r and r ^ 1 put their unit pair at the same bank group
    // ((u ^ (r & 7)) >> 1 is the same), so four CONSECUTIVE rows per quarter hit only 16 of the 32 banks twice each
    // (ncu: l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st = 58 % of the kernel's store wavefronts, the L1 data
    // pipe the busiest unit at 59 %).  Quarter q therefore owns rows 8 (q >> 1) + (q & 1) + {0, 2, 4, 6}: four different
    // bank groups, one wavefront per quarter.
    static __device__ __forceinline__ int tile_row(int ptid) {
        const int i = (ptid & 31) >> 1, q = i >> 2;
        return (ptid >> 5) * 16 + ((q >> 1) << 3) + ((i & 3) << 1) + (q & 1);
    }
    __device__ __forceinline__ void begin_tile(long long tile, int ptid) {
        const int r = tile_row(ptid);
        half = ptid & 1;
        long long row = tile * TILE_M + r;
        if (row >= g.m) row = g.m - 1;                       // padded rows recompute the last point; never stored
        const int bq = (int)(row / a.s);
        if (a.order != nullptr) row = (long long)bq * a.s + __ldg(a.order + (long long)bq * a.order_stride + (row - (long long)bq * a.s));
        boff = bq * a.n_cand;
        ip = a.idx + row * KN;
        const float *qp = a.query_xyz + row * 3;
        qx = qp[0]; qy = qp[1]; qz = qp[2];
    }

    __device__ __forceinline__ void load_idx(int pass, int (&gi)[NB]) {
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            gi[k] = boff + __ldg(ip + pass * NB + k);
            nb[k] = (uint32_t)gi[k] * (uint32_t)a.d;
        }
    }

    // This thread's 8 WeightNet outputs for the neighbours of one pass.  The relative coordinates are in registers
    // here anyway, so the pass's first K-chunk - channels (dx, dy, dz, 0) - is produced right here: the hot feature
    // loop carries no coordinate state and no branch.
    __device__ __forceinline__ void weightnet(const int (&gi)[NB], bool emit_xyz, unsigned char *a_hi, unsigned char *a_lo, int r) {
        float acc[3][8];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const float *cp = a.cand_xyz + (long long)gi[k] * 3;
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
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[0][j] = fmaf(dx, wn[k][j], acc[0][j]);
                acc[1][j] = fmaf(dy, wn[k][j], acc[1][j]);
                acc[2][j] = fmaf(dz, wn[k][j], acc[2][j]);
            }
        }
        if (emit_xyz) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
                if (c < 3) split8(acc[c < 3 ? c : 0], hi, lo);
                const uint32_t off = sw128_offset(r, c * 2 + half);
                *reinterpret_cast<uint4 *>(a_hi + off) = hi;
                *reinterpret_cast<uint4 *>(a_lo + off) = lo;
            }
        }
    }

    // Same with the WeightNet outputs read back from pointconv_weightnet_kernel's buffer: the unrolled 9-neighbour MLP
    // above is ~18 k cycles per tile inside this kernel (it spills around the wn[][] it is filling); precomputed, the
    // phase is 18 coalesced 16-byte loads and the coordinate chunk.
    __device__ __forceinline__ void weightnet_pre(const int (&gi)[NB], int pass, bool emit_xyz, unsigned char *a_hi,
                                                  unsigned char *a_lo, int r) {
        const float4 *wp = reinterpret_cast<const float4 *>(a.wn_pre + ((size_t)(ip - a.idx) + (size_t)pass * NB) * 16 + half * 8);
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const float4 lo4 = __ldg(wp + k * 4), hi4 = __ldg(wp + k * 4 + 1);
            wn[k][0] = lo4.x; wn[k][1] = lo4.y; wn[k][2] = lo4.z; wn[k][3] = lo4.w;
            wn[k][4] = hi4.x; wn[k][5] = hi4.y; wn[k][6] = hi4.z; wn[k][7] = hi4.w;
        }
        if (!emit_xyz) return;
        float acc[3][8];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const float *cp = a.cand_xyz + (long long)gi[k] * 3;
            const float dx = __ldg(cp) - qx, dy = __ldg(cp + 1) - qy, dz = __ldg(cp + 2) - qz;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[0][j] = fmaf(dx, wn[k][j], acc[0][j]);
                acc[1][j] = fmaf(dy, wn[k][j], acc[1][j]);
                acc[2][j] = fmaf(dz, wn[k][j], acc[2][j]);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
            if (c < 3) split8(acc[c < 3 ? c : 0], hi, lo);
            const uint32_t off = sw128_offset(r, c * 2 + half);
            *reinterpret_cast<uint4 *>(a_hi + off) = hi;
            *reinterpret_cast<uint4 *>(a_lo + off) = lo;
        }
    }

    __device__ __forceinline__ void gather(int fchunk, float4 (&v)[NB]) const {
        const float *fp = a.feats + fchunk * 4;
#pragma unroll
        for (int k = 0; k < NB; ++k) v[k] = __ldg(reinterpret_cast<const float4 *>(fp + nb[k]));
    }

    // one K-chunk (4 feature channels x this thread's 8 WeightNet outputs) from the gathered channels v[k]
    __device__ __forceinline__ void chunk_from(const float4 (&v)[NB], unsigned char *a_hi, int r) const {
        // packed fp32 FMAs (fma.rn.f32x2 -> SASS FFMA2): a 3-register FFMA issues every other cycle per sub-partition,
        // FFMA2 does two per issue; each half is an IEEE fma, so the sums are bit-identical to the scalar loop
        float2 acc2[4][4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[ch][j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const float2 vx = make_float2(v[k].x, v[k].x), vy = make_float2(v[k].y, v[k].y);
            const float2 vz = make_float2(v[k].z, v[k].z), vw = make_float2(v[k].w, v[k].w);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 w2 = make_float2(wn[k][2 * j], wn[k][2 * j + 1]);
                acc2[0][j] = __ffma2_rn(vx, w2, acc2[0][j]);
                acc2[1][j] = __ffma2_rn(vy, w2, acc2[1][j]);
                acc2[2][j] = __ffma2_rn(vz, w2, acc2[2][j]);
                acc2[3][j] = __ffma2_rn(vw, w2, acc2[3][j]);
            }
        }
        unsigned char *a_lo = a_hi + A_PART_BYTES;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint4 hi, lo;
            const float acc[8] = {acc2[ch][0].x, acc2[ch][0].y, acc2[ch][1].x, acc2[ch][1].y,
                                  acc2[ch][2].x, acc2[ch][2].y, acc2[ch][3].x, acc2[ch][3].y};
            split8(acc, hi, lo);
            const uint32_t off = sw128_offset(r, ch * 2 + half);
            *reinterpret_cast<uint4 *>(a_hi + off) = hi;
            *reinterpret_cast<uint4 *>(a_lo + off) = lo;
        }
    }

    template <class Acquire, class Release>
    __device__ __forceinline__ void run_tile(long long tile, int c_begin, int c_end, int ptid, unsigned char *, uint64_t *,
                                             Acquire &&acquire, Release &&release) {
        const int r = tile_row(ptid);
        begin_tile(tile, ptid);
        int c = c_begin;
        while (c < c_end) {
            const int pass = c / g.wchunks;
            const int base = pass * g.wchunks;
            const int stop = min(c_end, base + g.wchunks);
            int gi[NB];
            load_idx(pass, gi);
            if (c == base) {
                unsigned char *a_hi = acquire(c);
                if (a.wn_pre != nullptr) weightnet_pre(gi, pass, true, a_hi, a_hi + A_PART_BYTES, r);
                else weightnet(gi, true, a_hi, a_hi + A_PART_BYTES, r);
                release();
                ++c;
            } else {
                if (a.wn_pre != nullptr) weightnet_pre(gi, pass, false, nullptr, nullptr, r);
                else weightnet(gi, false, nullptr, nullptr, r);    // a split-K work item that starts inside a pass
            }
            // Feature chunks in PAIRS: the two half-lanes of a row fetch the two 16-byte halves of every neighbour's
            // 32-byte sector once (one L1 line access per row and pair instead of one per row and chunk) and hand each
            // other the half the current chunk needs by shuffle.
            const int lane = ptid & 31;
            if (c < stop && ((c - base - 1) & 1)) {            // (a split-K work item that starts on an odd feature chunk)
                unsigned char *a_hi = acquire(c);
                float4 v[NB];
                gather(c - base - 1, v);
                chunk_from(v, a_hi, r);
                release();
                ++c;
            }
#pragma unroll 1
            for (; c < stop; c += 2) {
                const int f0 = c - base - 1;                    // even feature chunk of the pair
                float4 mine[NB];
                if ((f0 + half) * 4 < a.d) {
                    gather(f0 + half, mine);
                } else {
#pragma unroll
                    for (int k = 0; k < NB; ++k) mine[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll 1
                for (int e = 0; e < 2; ++e) {
                    if (c + e >= stop) break;
                    unsigned char *a_hi = acquire(c + e);
                    const int src = (lane & ~1) | e;
                    float4 v[NB];
#pragma unroll
                    for (int k = 0; k < NB; ++k) {
                        v[k].x = __shfl_sync(0xffffffffu, mine[k].x, src);
                        v[k].y = __shfl_sync(0xffffffffu, mine[k].y, src);
                        v[k].z = __shfl_sync(0xffffffffu, mine[k].z, src);
                        v[k].w = __shfl_sync(0xffffffffu, mine[k].w, src);
                    }
                    chunk_from(v, a_hi, r);
                    release();
                }
            }
        }
    }
};

// WeightNet(3 -> 8 -> 8 -> 16, ReLU after every layer, pointconv_util.py:184-215) of every (query, neighbour) pair, one
// thread each, in exactly the fused kernel's operation order (bit-identical results): out [rows, KN, 16].
template <class Args>
__global__ void __launch_bounds__(256)
pointconv_weightnet_kernel(long long pairs, int kn, const Args a, float *__restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= pairs) return;
    const long long row = e / kn;
    const int b = (int)(row / a.s);
    const int nbr = b * a.n_cand + __ldg(a.idx + e);
    const float *qp = a.query_xyz + row * 3, *cp = a.cand_xyz + (long long)nbr * 3;
    const float dx = __ldg(cp) - __ldg(qp), dy = __ldg(cp + 1) - __ldg(qp + 1), dz = __ldg(cp + 2) - __ldg(qp + 2);
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
    float w[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        float t = a.b3[o];
#pragma unroll
        for (int i = 0; i < 8; ++i) t += a.w3[o * 8 + i] * h2[i];
        w[o] = fmaxf(t, 0.f);
    }
    float4 *o4 = reinterpret_cast<float4 *>(out + e * 16);
#pragma unroll
    for (int o = 0; o < 16; o += 4) o4[o >> 2] = make_float4(w[o], w[o + 1], w[o + 2], w[o + 3]);
}

static int kdpc_pointconv_stages = 2;
static int kdpc_pointconv_precompute = 1;                // 1 = where it pays (below), 0 = never, 2 = always (A/B measurements)

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
    // enough to keep the MMA fed and leave ~80 KB of the unified L1/shared array to the cache
    if (g.stages > kdpc_pointconv_stages) g.stages = kdpc_pointconv_stages;
    const size_t smem = smem_bytes(g.n_pad, g.stages);
    auto kern = tc_gemm_kernel<P, StoreEpilogue>;
    KDPC_ENSURE_SMEM(kern, 201 * 1024);
    const long long work = g.num_tiles * g.splits;
    const unsigned grid = (unsigned)(work < num_sms() ? work : num_sms());
    launch_tc(kern, grid, num_threads<P>(), smem, st, g, pa, ea);
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
/* A/B switch for measurements: 0 = WeightNet evaluated inside the fused kernel instead of by the small pre-pass (same results) */
KDPC_API void kdpc_pointconv_set_precompute(int on) { kdpc::tc::kdpc_pointconv_precompute = on; }

KDPC_API long long kdpc_pointconv_fused_ws_bytes(int b, int s, int k, int d, int n_out) {
    if (b <= 0 || s <= 0 || d <= 0 || n_out <= 0 || n_out > 256) return 0;
    // split-K partial sums (16-byte multiple) followed by the precomputed WeightNet outputs [b*s, k, 16]
    const long long splitk = ((long long)split_k_ws_bytes(pointconv_shape((long long)b * s, n_out, d, k)) + 15) / 16 * 16;
    return splitk + (long long)b * s * k * 16 * (long long)sizeof(float);
}

KDPC_API int kdpc_pointconv_fused_ordered(int b, int n, int s, int k, int d, int n_out, const float *cand_xyz,
                                          const float *query_xyz, const float *feats, const int *idx,
                                          const float *wn_params /* host: w1[24] b1[8] w2[64] b2[8] w3[128] b3[16] */,
                                          const void *wpacked, const float *scale, const float *shift, float slope,
                                          const int *row_order, int order_stride, void *ws, float *out,
                                          kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(cand_xyz && query_xyz && feats && idx && wn_params && wpacked && out && b > 0 && n > 0 && s > 0 &&
                    d > 0 && n_out > 0);
    if (n_out > 256 || (d & 3) != 0 || (k != 9 && k != 16)) return KDPC_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(feats) % 16) != 0 || (reinterpret_cast<uintptr_t>(out) % 16) != 0 ||
        (reinterpret_cast<uintptr_t>(wpacked) % 16) != 0 || (reinterpret_cast<uintptr_t>(ws) % 16) != 0)
        return KDPC_EINVAL;
    if (row_order != nullptr && order_stride < s) return KDPC_EINVAL;
    PointConvProducer<9, 1>::Args pa;                 // (the Args layout does not depend on the template arguments)
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
    // the row order only pays (and is only supported) without split-K: small layers keep the natural order
    const GemmShape planned = pointconv_shape((long long)b * s, n_out, d, k);
    pa.wn_pre = nullptr;
    // the pre-pass pays where the in-kernel evaluation is expensive: many tiles per SM (the 18 k-cycle phase serialises
    // with the chunk loop), or split-K (every work item of a tile would evaluate the same WeightNet again); measured
    // slower for the layers in between (an extra launch for ~3 k rows per SM)
    const bool pre = kdpc_pointconv_precompute == 2 || (kdpc_pointconv_precompute == 1 && (planned.splits > 1 || (long long)b * s >= 32768));
    if (ws != nullptr && pre) {
        float *wn_ws = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(ws) + (split_k_ws_bytes(planned) + 15) / 16 * 16);
        const long long pairs = (long long)b * s * k;
        pointconv_weightnet_kernel<<<(unsigned)div_up_ll(pairs, 256), 256, 0, to_stream(stream)>>>(pairs, k, pa, wn_ws);
        pa.wn_pre = wn_ws;
    }
    void *splitk_ws = planned.splits > 1 ? ws : nullptr;     // (the launcher plans split-K only when it gets a workspace)
    const bool ordered = row_order != nullptr && planned.splits == 1;
    pa.order = ordered ? row_order : nullptr;
    pa.order_stride = ordered ? order_stride : 0;
    if (ordered) { ea.row_order = row_order; ea.order_stride = order_stride; ea.rows_per_cloud = s; }
    if (k == 9) return launch_pointconv<9, 1>((long long)b * s, n_out, pa, wpacked, ea, splitk_ws, to_stream(stream));
    PointConvProducer<16, 2>::Args pb;
    static_assert(sizeof(pb) == sizeof(pa), "Args layout");
    memcpy(&pb, &pa, sizeof(pa));
    return launch_pointconv<16, 2>((long long)b * s, n_out, pb, wpacked, ea, splitk_ws, to_stream(stream));
}

KDPC_API int kdpc_pointconv_fused(int b, int n, int s, int k, int d, int n_out, const float *cand_xyz,
                                  const float *query_xyz, const float *feats, const int *idx, const float *wn_params,
                                  const void *wpacked, const float *scale, const float *shift, float slope,
                                  void *ws, float *out, kdpc_stream_t stream) {
    return kdpc_pointconv_fused_ordered(b, n, s, k, d, n_out, cand_xyz, query_xyz, feats, idx, wn_params, wpacked, scale,
                                        shift, slope, nullptr, 0, ws, out, stream);
}
