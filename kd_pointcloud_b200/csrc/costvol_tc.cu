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
    static constexpr bool kAsync = false;
    static constexpr int kIssuers = 0, kLookahead = 0;
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

// Asynchronous variant.  The positional encoding is linear, pos_w (xyz2[j] - xyz1[i]) + pos_b =
// (pos_w xyz2[j]) - (pos_w xyz1[i]) + pos_b, so a tiny elementwise pass (costvol_prep_kernel) folds it into the
// point features once per POINT instead of once per (point, neighbour, channel):
//     p2q[j] = points2[j] + pos_w xyz2[j]          p1q[i] = points1[i] + pos_b - pos_w xyz1[i]
// and a row of the A operand is just act(p2q[idx] + p1q[i]).  Everything a row needs arrives through cp.async
// (LDGSTS, 16-byte pieces; per-row bulk copies of 128 B turned out to be TMA-issue bound) two pipeline
// iterations ahead of its conversion, so the converting threads only ever read shared memory; only the
// neighbour index travels in a register (loaded one iteration before it is needed).  Completion:
// cp.async.mbarrier.arrive on the raw stage's mbarrier, one arrival per producer thread.
// 8 producer warps: thread (row, half) converts half of the chunk's channels.
struct CostVolAsyncProducer {
    static constexpr int kWarps = 8, kGroups = 1;
    static constexpr bool kAsync = true;
    // 4 extra warps do nothing but issue the gathers (a warp-level LDGSTS costs ~28 cycles of load/store-unit time and
    // blocks its warp meanwhile: issued by the converting warps it took 1000-1350 of the 2250-2650 cycles of a pipeline
    // iteration, tools/trace_costvol.py); the 8 producer warps only convert
#ifndef KDPC_CV_IW
#define KDPC_CV_IW 4
#endif
    static constexpr int kIssuerWarps = KDPC_CV_IW;
    static constexpr int RSTEP = 32 * kIssuerWarps / 8;          // thread t serves rows (t >> 3) + RSTEP j
    static constexpr int JPP = CV_K / RSTEP;                     // consecutive j that belong to one point
    static constexpr int kIssuers = 32 * kIssuerWarps, kLookahead = 2;
#ifndef KDPC_CV_SWIZZLE
#define KDPC_CV_SWIZZLE 1
#endif
    // Raw staging rows.  KDPC_CV_SWIZZLE = 1 (default): 256-byte rows, the eight 16-byte pieces of each 128-byte half stored
    // at piece ^ (row & 7) - every LDGSTS quarter warp writes ONE aligned 128-byte line, and the converters' 16-byte reads
    // by row stay conflict-free.  0: the padded layout (256 B + 16 B) it replaces: its rows start 16 bytes off the
    // 32-byte sectors on every other row and ncu counted 10.5 shared-memory wavefronts per warp-level LDGSTS against
    // an ideal of 4 (L1 Wavefronts Shared Excessive = 62 %; the L1 data pipe is the cost volume's busiest unit).
    static constexpr int ROW_PITCH = KDPC_CV_SWIZZLE ? 256 : 272;
    static constexpr int kRawBytes = (TILE_M + TILE_M / CV_K) * ROW_PITCH;     // 128 neighbour rows + 4 point rows
    static constexpr int RPT = TILE_M * 8 / kIssuers;        // neighbour rows per issuing thread (8 lanes per row)
    struct Args {
        const float *p1q;    // [B,S,D]  points1 + pos_b - pos_w xyz1
        const float *p2q;    // [B,N,D]  points2 + pos_w xyz2
        const int *idx;      // [B,S,32]
        int s, n, d;
        float slope;
        long long points;    // B * S (the paired producer's row bound; its GemmShape counts 128-row iterations)
    };
    static __device__ __forceinline__ void prologue(const Args &, int, int) {}
    const Args &a;
    const GemmShape &g;
    // (Measured and rejected: L1-allocating cp.async.ca with the queries in Morton order - no gain; the cost is per
    // LDGSTS instruction, ~28 cycles of load/store-unit time each, whatever the hit rate.)
    // Gather issue mapping: 8 consecutive lanes copy the (up to two) 128-byte halves of ONE neighbour row's chunk slice,
    // 16 bytes each, so a warp-level LDGSTS touches 4 lines; issuing thread t serves rows (t >> 3) + 16 j, j = 0..7.  (With one
    // thread per row copying its row's pieces one after the other every request touched 32 lines and the load/store
    // unit needed ~1350 of the 2650 cycles of a pipeline iteration just to accept them - tools/trace_costvol.py.)
    // All per-row address arithmetic happens once per tile, one iteration ahead (load_rows): the issue itself is an
    // add and a predicated LDGSTS per row.
    int nbr[RPT];                 // neighbour index of tile row (t >> 3) + 16 j in the NEXT issued tile (loaded one iteration ahead,
                                  // first used at that issue: the load latency never stalls the issuing warp)
    uint32_t cloud_off[RPT];      // (cloud of that row) * n
    uint32_t pt_off;              // element offset into p1q of the point row this thread copies a piece of
    uint32_t dst0;                // byte offset of this thread's piece inside a raw staging buffer (row j = 0)

    __device__ CostVolAsyncProducer(const Args &a_, const GemmShape &g_) : a(a_), g(g_) {}

    // rows < 2^31, p1q / p2q elements < 2^32 (checked on the host): all index arithmetic in 32 bits
    __device__ __forceinline__ unsigned row_of(int tile, int r) const {
        const unsigned row = (unsigned)tile * TILE_M + (unsigned)r;
        return row < (unsigned)g.m ? row : (unsigned)g.m - 1u;   // padded rows repeat the last row; never stored
    }
    __device__ __forceinline__ void load_rows(int tile, int ptid) {
        const unsigned s = (unsigned)a.s, pt0 = (unsigned)tile * (TILE_M / CV_K);
        const unsigned b0 = pt0 / s, left = (b0 + 1u) * s - pt0;  // points of the tile before the next cloud starts
        const unsigned last_pt = (unsigned)(g.m >> 5) - 1u;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {                           // tile row (t >> 3) + RSTEP j belongs to point pt0 + j / JPP
            const unsigned jj = min((unsigned)(j / JPP), last_pt - pt0);  // (rows past the end repeat the last point's last row)
            const unsigned b = b0 + (jj >= left ? (jj - left) / s + 1u : 0u);
            nbr[j] = __ldg(a.idx + row_of(tile, (ptid >> 3) + RSTEP * j));
            cloud_off[j] = b * (unsigned)a.n;
        }
        const unsigned pt = min(pt0 + (unsigned)((ptid >> 5) & 3), last_pt);
        pt_off = pt * (unsigned)a.d;
    }
    __device__ __forceinline__ void prime(int tile, int ptid) {
        static_assert(RSTEP % 8 == 0, "the swizzle term (row & 7) must not depend on j");
        dst0 = (uint32_t)((ptid >> 3) * ROW_PITCH + ((KDPC_CV_SWIZZLE ? ((ptid & 7) ^ ((ptid >> 3) & 7)) : (ptid & 7)) << 4));
        load_rows(tile, ptid);
    }

    // channels [c0, c0 + 64) of a chunk are split between the two half-threads of a row: units of 8 channels,
    // half 0 takes the first ceil(units/2)
    __device__ __forceinline__ void split_units(int chunk, int half, int &u0, int &nu) const {
        const int units = min(8, (a.d - chunk * CHUNK_K) >> 3);
        const int first = (units + 1) >> 1;
        u0 = half ? first : 0;
        nu = half ? units - first : first;
    }

    __device__ __forceinline__ void issue(int /*tile*/, int chunk, int next_tile, unsigned char *raw, uint64_t *bar, int ptid) {
        const int q = ptid & 7;
        const int c0 = chunk * CHUNK_K;
        const int ppr = min(CHUNK_K, a.d - c0) >> 2;             // 16-byte pieces per row in this chunk (d % 8 == 0)
        const uint32_t dst = smem_u32(raw) + dst0;
        const float *src = a.p2q + c0 + q * 4;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const uint32_t row_off = (cloud_off[j] + (uint32_t)nbr[j]) * (uint32_t)a.d;
            if (q < ppr) cp_async_16(dst + j * (RSTEP * ROW_PITCH), src + row_off);
            if (q + 8 < ppr) cp_async_16(dst + j * (RSTEP * ROW_PITCH) + 128, src + row_off + 32);
        }
        {                                                        // the 4 points' own rows, one 16-byte piece per lane
            const int k = ptid & 31;
            if (k < ppr && ptid < 128)
                cp_async_16(smem_u32(raw + (TILE_M + ((ptid >> 5) & 3)) * ROW_PITCH) + k * 16, a.p1q + pt_off + c0 + k * 4);
        }
        cp_async_mbar_arrive(bar);                               // arrives once this thread's copies have landed
        if (next_tile >= 0) load_rows(next_tile, ptid);          // consumed by the next issue
    }

    __device__ __forceinline__ void convert(int tile, int chunk, const unsigned char *raw, unsigned char *a_hi,
                                            unsigned char *a_lo, int ptid) {
        const int r = ptid & 127, half = ptid >> 7;
        const float4 *rr = reinterpret_cast<const float4 *>(raw + r * ROW_PITCH);
        const float4 *pr = reinterpret_cast<const float4 *>(raw + (TILE_M + (r >> 5)) * ROW_PITCH);   // same for the warp
        int u0, nu;
        split_units(chunk, half, u0, nu);
        const float slope = a.slope;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
            if (uu < nu) {
                const int u = u0 + uu;
                const int sw = KDPC_CV_SWIZZLE ? (r & 7) : 0;                          // pieces of a 128-byte half: piece ^ (row & 7)
                const float4 g0 = rr[(2 * u) ^ sw], g1 = rr[(2 * u + 1) ^ sw];
                const float4 q0 = pr[2 * u], q1 = pr[2 * u + 1];
                float v[8] = {g0.x + q0.x, g0.y + q0.y, g0.z + q0.z, g0.w + q0.w,
                              g1.x + q1.x, g1.y + q1.y, g1.z + q1.z, g1.w + q1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], v[e] * slope);          // leaky / ReLU (0 <= slope < 1)
                uint4 hi, lo;
                split8(v, hi, lo);
                const uint32_t off = sw128_offset(r, u);
                *reinterpret_cast<uint4 *>(a_hi + off) = hi;
                *reinterpret_cast<uint4 *>(a_lo + off) = lo;
            }
        }
        // units beyond d (d < 64): the MMA never reads them (k_total stops the K loop at d rounded up to 16)
        if (half == 1) {
            const int units = min(8, (a.d - chunk * CHUNK_K) >> 3);
            if (units & 1) {                                      // ... except the upper half of an odd last 16-wide K-step
                const uint32_t off = sw128_offset(r, units);
                *reinterpret_cast<uint4 *>(a_hi + off) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(a_lo + off) = make_uint4(0, 0, 0, 0);
            }
        }
    }
};

// PAIRED variant for the 8192-point level (D = 32, D' <= 32).  A 128-row x K=32 tile is too little work per pipeline
// iteration: tools/trace_costvol.py shows the single MMA warp's per-iteration control path (three mbarrier waits, the
// fence, descriptors, two commits: ~1750 cycles) pacing the kernel while the tensor pipe is ~5 % busy.  So one
// iteration carries TWO row tiles (8 points): the A stage row r holds [tile 0 row r, channels 0..31 | tile 1 row r,
// channels 0..31] (K' = 64) and the weight is the block-diagonal 64 x 64 matrix diag(W, W) (costvol_pair_weight_kernel),
// so the accumulator's columns 0..31 are tile 0's outputs and columns 32..63 tile 1's.  The tensor pipe multiplies the
// two zero blocks as well (it has the time); waits, commits, barriers and loop overhead are paid once per 256 rows by
// every role, and EIGHT epilogue warps (two per TMEM lane quarter, one per column half) keep the max-over-K off the
// critical path.  126 -> 101-105 us at B = 8 x 8192 points, bit-identical.  What paces it now is the load/store unit: 66
// warp-level LDGSTS per iteration (~28 cycles each) plus the conversion's LDS/STS (tools/trace_costvol.py).
// Measured and rejected for the gathers (tools/gather4_probe.cu, same box): cp.async.bulk.tensor tile::gather4 (one
// TMA request per 4 rows) - the TMA unit takes ~150 cycles per request, 5x the LDGSTS cost per row; plain LDG.128 ->
// registers -> STS.128 from the issuer warps - halves the conversion time (no LDGSTS in the LSU queue) but one
// iteration of loads in flight per warp cannot cover the L2 latency (158 us).  8 issuer warps instead of 4: -5 % here (kept),
// +3..5 % for the unpaired producer at D = 64 / 256 (not adopted there).
struct CostVolPairProducer {
    static constexpr int kWarps = 8, kGroups = 1;
    static constexpr bool kAsync = true;
#ifndef KDPC_CV_PAIR_IW
#define KDPC_CV_PAIR_IW 8
#endif
    static constexpr int kIssuerWarps = KDPC_CV_PAIR_IW, kEpilogueWarps = 8;
#ifndef KDPC_CV_RELAXED
#define KDPC_CV_RELAXED 1
#endif
    static constexpr bool kRelaxedWaits = KDPC_CV_RELAXED != 0;          // nanosleep back-off in the converters' / epilogue warps' waits (tc_gemm.cuh)
    static constexpr int kIssuers = 32 * kIssuerWarps, kLookahead = 2;
    static constexpr int D = 32;
    static constexpr int ROWS = 2 * TILE_M, PTS = ROWS / CV_K;           // 256 neighbour rows = 8 points per iteration
    static constexpr int ROW_PITCH = KDPC_CV_SWIZZLE ? 128 : 144;        // swizzled 128-byte rows (see CostVolAsyncProducer), or 128 B + 16 B padding
    static constexpr int kRawBytes = (ROWS + PTS) * ROW_PITCH;
    static constexpr int RPT = ROWS * 8 / kIssuers;                      // neighbour rows per issuing thread (8 lanes per row)
    static constexpr int RSTEP = kIssuers / 8;                           // thread t serves rows (t >> 3) + RSTEP j
    static constexpr int JPP = CV_K / RSTEP;                             // consecutive j that belong to one point
    using Args = CostVolAsyncProducer::Args;
    static __device__ __forceinline__ void prologue(const Args &, int, int) {}
    const Args &a;
    const GemmShape &g;
    uint32_t roff[RPT];           // element offset into p2q of the neighbour row that lands in rows (t >> 3) + RSTEP j of the NEXT issued iteration
    uint32_t pt_off;              // element offset into p1q of the point row this thread copies a piece of (threads < 64)
    uint32_t dst0;
    uint32_t last_pt, last_row;

    __device__ CostVolPairProducer(const Args &a_, const GemmShape &g_) : a(a_), g(g_) {}

    // g.m = 128 x iterations ("virtual" rows); the real row count is points * 32 with points = a.points
    __device__ __forceinline__ void load_rows(int tile, int ptid) {
        const unsigned s = (unsigned)a.s, pt0 = (unsigned)tile * PTS;
        const unsigned b0 = pt0 / s, left = (b0 + 1u) * s - pt0;  // points of the iteration before the next cloud starts
#pragma unroll
        for (int j = 0; j < RPT; ++j) {                           // row (t >> 3) + RSTEP j belongs to point pt0 + j / JPP
            const unsigned jj = min((unsigned)(j / JPP), last_pt - pt0);   // (rows past the end repeat the last point's last row)
            const unsigned b = b0 + (jj >= left ? (jj - left) / s + 1u : 0u);
            const unsigned row = pt0 * CV_K + (unsigned)((ptid >> 3) + RSTEP * j);
            roff[j] = (b * (unsigned)a.n + (unsigned)__ldg(a.idx + min(row, last_row))) * (unsigned)D;
        }
        const unsigned pt = min(pt0 + (unsigned)((ptid >> 3) & 7), last_pt);
        pt_off = pt * (unsigned)D;
    }
    __device__ __forceinline__ void prime(int tile, int ptid) {
        static_assert(RSTEP % 8 == 0, "the swizzle term (row & 7) must not depend on j");
        dst0 = (uint32_t)((ptid >> 3) * ROW_PITCH + ((KDPC_CV_SWIZZLE ? ((ptid & 7) ^ ((ptid >> 3) & 7)) : (ptid & 7)) << 4));
        last_pt = (unsigned)a.points - 1u;
        last_row = (unsigned)a.points * CV_K - 1u;
        load_rows(tile, ptid);
    }
    __device__ __forceinline__ void issue(int /*tile*/, int /*chunk*/, int next_tile, unsigned char *raw, uint64_t *bar, int ptid) {
        const float *src = a.p2q + (ptid & 7) * 4;
        const uint32_t dst = smem_u32(raw) + dst0;
#pragma unroll
        for (int j = 0; j < RPT; ++j)
            cp_async_16(dst + j * (RSTEP * ROW_PITCH), src + roff[j]);
        if (ptid < 8 * PTS)                                       // the 8 points' own rows: 8 pieces each
            cp_async_16(smem_u32(raw + (ROWS + (ptid >> 3)) * ROW_PITCH) + (ptid & 7) * 16, a.p1q + pt_off + (ptid & 7) * 4);
        cp_async_mbar_arrive(bar);
        if (next_tile >= 0) load_rows(next_tile, ptid);
    }
    // thread (r, half): row r of row tile `half`, all 32 channels -> units 4 half .. 4 half + 3 of stage row r
    __device__ __forceinline__ void convert(int /*tile*/, int /*chunk*/, const unsigned char *raw, unsigned char *a_hi,
                                            unsigned char *a_lo, int ptid) {
        const int r = ptid & 127, half = ptid >> 7;
        const float4 *rr = reinterpret_cast<const float4 *>(raw + (half * TILE_M + r) * ROW_PITCH);
        const float4 *pr = reinterpret_cast<const float4 *>(raw + (ROWS + half * (TILE_M / CV_K) + (r >> 5)) * ROW_PITCH);   // same for the warp
        const float slope = a.slope;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int sw = KDPC_CV_SWIZZLE ? (r & 7) : 0;
            const float4 g0 = rr[(2 * u) ^ sw], g1 = rr[(2 * u + 1) ^ sw];
            const float4 q0 = pr[2 * u], q1 = pr[2 * u + 1];
            float v[8] = {g0.x + q0.x, g0.y + q0.y, g0.z + q0.z, g0.w + q0.w,
                          g1.x + q1.x, g1.y + q1.y, g1.z + q1.z, g1.w + q1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], v[e] * slope);          // leaky / ReLU (0 <= slope < 1)
            uint4 hi, lo;
            split8(v, hi, lo);
            const uint32_t off = sw128_offset(r, half * 4 + u);
            *reinterpret_cast<uint4 *>(a_hi + off) = hi;
            *reinterpret_cast<uint4 *>(a_lo + off) = lo;
        }
    }
};

// diag(W, W) in the packed layout (one chunk, n_pad = 64) from the packed W (one chunk, n_pad_src = 32 or 16):
// rows 0..31 keep W's units 0..3, rows 32..63 carry them as units 4..7, everything else is zero
__global__ void __launch_bounds__(256)
costvol_pair_weight_kernel(int n_pad_src, const unsigned char *__restrict__ src, unsigned char *__restrict__ dst) {
    const int t = threadIdx.x + blockIdx.x * blockDim.x;           // (part, row, unit): 2 x 64 x 8
    if (t >= 2 * 64 * 8) return;
    const int u = t & 7, row = (t >> 3) & 63, part = t >> 9;
    const int srow = row & 31, su = row < 32 ? u : u - 4;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (su >= 0 && su < 4 && srow < n_pad_src)
        v = *reinterpret_cast<const uint4 *>(src + (size_t)part * n_pad_src * 128 + sw128_offset(srow, su));
    *reinterpret_cast<uint4 *>(dst + (size_t)part * 64 * 128 + sw128_offset(row, u)) = v;
}

// p2q = points2 + pos_w xyz2 (sign = +1, no bias);  p1q = points1 + pos_b - pos_w xyz1 (sign = -1, with bias)
__global__ void __launch_bounds__(256)
costvol_prep_kernel(long long rows1, long long rows2, int dvec, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                    const float *__restrict__ p1, const float *__restrict__ p2, const float *__restrict__ pos_w,
                    const float *__restrict__ pos_b, float *__restrict__ p1q, float *__restrict__ p2q) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n1 = rows1 * dvec, n2 = rows2 * dvec;
    if (e >= n1 + n2) return;
    const bool first = e < n1;
    const long long ee = first ? e : e - n1;
    const long long row = ee / dvec;
    const int dv = (int)(ee - row * dvec);
    const float *x = (first ? xyz1 : xyz2) + row * 3;
    const float sx = first ? -x[0] : x[0], sy = first ? -x[1] : x[1], sz = first ? -x[2] : x[2];
    const float4 v = __ldg(reinterpret_cast<const float4 *>(first ? p1 : p2) + ee);
    const float4 bb = first ? __ldg(reinterpret_cast<const float4 *>(pos_b) + dv) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float *w = pos_w + (size_t)dv * 12;
    float4 o;
    o.x = v.x + (bb.x + w[0] * sx + w[1] * sy + w[2] * sz);
    o.y = v.y + (bb.y + w[3] * sx + w[4] * sy + w[5] * sz);
    o.z = v.z + (bb.z + w[6] * sx + w[7] * sy + w[8] * sz);
    o.w = v.w + (bb.w + w[9] * sx + w[10] * sy + w[11] * sz);
    reinterpret_cast<float4 *>(first ? p1q : p2q)[ee] = o;
}

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
        int pair_n;           // 0, or D' of the paired layout: accumulator columns 32 h .. 32 h + D' - 1 = row tile h of the iteration
    };
    __device__ __forceinline__ void tile(const Args &e, const GemmShape &g, long long tile, int split, uint32_t t_acc,
                                         int quarter, int lane) const {
        if (e.pair_n > 0) {                                       // eight warps: (quarter & 3) = point of the row tile, quarter >> 2 = row tile
            const int h = quarter >> 2;
            pair_tile(e, tile * (2 * TILE_M / CV_K) + h * (TILE_M / CV_K) + (quarter & 3), t_acc + (uint32_t)(32 * h), lane);
            return;
        }
        const long long pt = tile * (TILE_M / CV_K) + quarter;
        const int n0 = g.nsplit ? split * g.n_pad : 0;            // split-N: this work item's first output column
        for (int c0 = 0; c0 < g.n_pad; c0 += 32) {
            float v[32];
            tmem_ld_32x32(t_acc + (uint32_t)c0, v);               // lane = neighbour k, v[j] = channel c0+j
            // max over the 32 lanes (neighbours) of every column, lane j ending up with column j: recursive halving -
            // at distance 16, 8, .. 1 every lane keeps the half of its columns that matches its lane bit and trades the
            // other half with its partner: 31 SHFL + 31 integer max.  (32 x redux.sync.max was ~1800 cycles per tile:
            // the REDUX results come back through the uniform datapath one at a time, ~56 cycles each.)
            int w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = f2ord(v[j]);
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const bool up = (lane & m) != 0;
#pragma unroll
                for (int j = 0; j < m; ++j) {
                    const int keep = up ? w[j + m] : w[j];
                    const int give = up ? w[j] : w[j + m];
                    w[j] = max(keep, __shfl_xor_sync(0xffffffffu, give, m));
                }
            }
            const int mine = w[0];
            const int col = n0 + c0 + lane;
            if (pt < e.points && col < g.n) {
                float y = ord2f(mine);
                if (e.bias) y += __ldg(e.bias + col);
                y = y > 0.f ? y : y * e.slope;
                e.out[pt * e.ldo + col] = y;
            }
        }
    }
    // one point's 32 neighbours x (up to) 32 channels at t_acc: same transpose-reduce
    __device__ __forceinline__ void pair_tile(const Args &e, long long pt, uint32_t t_acc, int lane) const {
        float v[32];
        tmem_ld_32x32(t_acc, v);
        int w[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) w[j] = f2ord(v[j]);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const bool up = (lane & m) != 0;
#pragma unroll
            for (int j = 0; j < m; ++j) {
                const int keep = up ? w[j + m] : w[j];
                const int give = up ? w[j] : w[j + m];
                w[j] = max(keep, __shfl_xor_sync(0xffffffffu, give, m));
            }
        }
        if (pt < e.points && lane < e.pair_n) {
            float y = ord2f(w[0]);
            if (e.bias) y += __ldg(e.bias + lane);
            y = y > 0.f ? y : y * e.slope;
            e.out[pt * e.ldo + lane] = y;
        }
    }
};

}  // namespace tc
}  // namespace kdpc

using namespace kdpc;
using namespace kdpc::tc;

static int kdpc_costvol_colblocks = 1;
static int kdpc_costvol_pairing = 1;
/* A/B switch for measurements: 0 = one 128-row tile per pipeline iteration at every level (same results) */
KDPC_API void kdpc_costvol_set_pairing(int on) { kdpc_costvol_pairing = on & 1; kdpc_costvol_colblocks = !(on & 2); }   // bit 1: no column blocks

KDPC_API long long kdpc_costvol_fused_ws_bytes(int b, int s, int n, int d) {
    if (b <= 0 || s <= 0 || n <= 0 || d <= 0) return 0;
    return ((long long)b * s + (long long)b * n) * d * 4 + 2 * 64 * 128;     // p1q, p2q, the paired layout's diag(W, W)
}

KDPC_API int kdpc_costvol_fused(int b, int s, int n, int k, int d, int d_out, const float *xyz1, const float *xyz2,
                                const float *p1, const float *p2, const int *idx, const float *pos_w,
                                const float *pos_b, float slope_pre, const void *wpacked, const float *bias,
                                float slope_post, void *ws, float *out, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(xyz1 && xyz2 && p1 && p2 && idx && pos_w && pos_b && wpacked && out && b > 0 && s > 0 && n > 0 &&
                    d > 0 && d_out > 0);
    if (k != CV_K || d > CV_MAX_D || (d & 7) != 0 || d_out > 256 || (long long)b * s * CV_K >= (1ll << 31) ||
        (long long)b * n * d >= (1ll << 32) || (long long)b * s * d >= (1ll << 32)) return KDPC_EUNSUPPORTED;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2) | reinterpret_cast<uintptr_t>(wpacked) |
                         reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(pos_b);
    if (al % 16 != 0) return KDPC_EINVAL;
    const long long points = (long long)b * s;
    MaxKEpilogue::Args ea{bias, slope_post, out, d_out, points, 0};
    if (ws != nullptr && slope_pre >= 0.f && slope_pre < 1.f && kdpc_tc_async_enabled() && kdpc_costvol_pairing &&
        d == CostVolPairProducer::D && d_out > 16 && d_out <= 32 && points >= 4096) {
        // 8192-point level: two row tiles per pipeline iteration against diag(W, W)
        using P = CostVolPairProducer;
        const long long iters = (points + P::PTS - 1) / P::PTS;
        float *p1q = reinterpret_cast<float *>(ws);
        float *p2q = p1q + points * d;
        unsigned char *w2 = reinterpret_cast<unsigned char *>(p2q + (long long)b * n * d);
        GemmShape g = make_shape(iters * TILE_M, 32 + d_out, 2 * d, w2, P::kRawBytes, P::kLookahead + 1);
        g.stages = 2;             // 2 x 48 KB of operand stages + 3 x 37 KB of staging: 212 KB (nothing here lives on L1 hits)
        if (g.n_pad == 64) {
            const long long total = (points + (long long)b * n) * (d / 4);
            costvol_prep_kernel<<<(unsigned)div_up_ll(total, 256), 256, 0, to_stream(stream)>>>(
                points, (long long)b * n, d / 4, xyz1, xyz2, p1, p2, pos_w, pos_b, p1q, p2q);
            costvol_pair_weight_kernel<<<4, 256, 0, to_stream(stream)>>>((d_out + 15) / 16 * 16, reinterpret_cast<const unsigned char *>(wpacked), w2);
            P::Args pa{p1q, p2q, idx, s, n, d, slope_pre, points};
            ea.pair_n = d_out;
            const size_t smem = smem_bytes(g.n_pad, g.stages, g.raw_bytes * g.raw_stages);
            auto kern = tc_gemm_kernel<P, MaxKEpilogue>;
            KDPC_ENSURE_SMEM(kern, 216 * 1024);
            const unsigned grid = (unsigned)(g.num_tiles < num_sms() ? g.num_tiles : num_sms());
            launch_tc(kern, grid, num_threads<P>(), smem, to_stream(stream), g, pa, ea);
            KDPC_RETURN_LAST();
        }
    }
    if (ws != nullptr && slope_pre >= 0.f && slope_pre < 1.f && kdpc_tc_async_enabled()) {
        // asynchronous producer whenever its raw staging fits next to >= 2 operand stages
        using P = CostVolAsyncProducer;
        GemmShape g = make_shape(points * CV_K, d_out, d, wpacked, P::kRawBytes, P::kLookahead + 1);
        if (g.stages < 2) g = make_shape(points * CV_K, d_out, d, wpacked, P::kRawBytes, P::kLookahead);   // lookahead 1
        if (g.stages < 2 && kdpc_costvol_colblocks && g.n_pad > 128 && (g.n_pad & 127) == 0) {
            // D' = 256 (the 256-point level): 64 KB weight stages leave no room for the staging ring.  Work items become
            // (row tile, 128-column block): 32 KB weight stages, the asynchronous producer fits - at the price of building
            // every A tile once per column block (the synchronous register-staged producer was 72 us per call for 2048 points)
            g.w_n_pad = g.n_pad;
            g.splits = g.n_pad / 128;
            g.nsplit = 1;
            g.n_pad = 128;
            g.acc_stride = 128;
            g.nacc_log2 = 2;
            g.tmem_cols = 512;
            g.stages = 2;                  // 2 x 64 KB of operand stages + 2 x 36 KB of staging: 201 KB
        }
        if (g.stages >= 2) {
            float *p1q = reinterpret_cast<float *>(ws);
            float *p2q = p1q + points * d;
            const long long total = (points + (long long)b * n) * (d / 4);
            costvol_prep_kernel<<<(unsigned)div_up_ll(total, 256), 256, 0, to_stream(stream)>>>(
                points, (long long)b * n, d / 4, xyz1, xyz2, p1, p2, pos_w, pos_b, p1q, p2q);
            P::Args pa{p1q, p2q, idx, s, n, d, slope_pre, points};
            const size_t smem = smem_bytes(g.n_pad, g.stages, g.raw_bytes * g.raw_stages);
            auto kern = tc_gemm_kernel<P, MaxKEpilogue>;
            KDPC_ENSURE_SMEM(kern, 216 * 1024);
            const long long work = g.num_tiles * (g.nsplit ? g.splits : 1);
            const unsigned grid = (unsigned)(work < num_sms() ? work : num_sms());
            launch_tc(kern, grid, num_threads<P>(), smem, to_stream(stream), g, pa, ea);
            KDPC_RETURN_LAST();
        }
    }
    CostVolProducer::Args pa{xyz1, xyz2, p1, p2, idx, pos_w, pos_b, s, n, d, slope_pre};
    GemmShape g = make_shape(points * CV_K, d_out, d, wpacked);
    const size_t smem = smem_bytes(g.n_pad, g.stages);
    auto kern = tc_gemm_kernel<CostVolProducer, MaxKEpilogue>;
    KDPC_ENSURE_SMEM(kern, 201 * 1024);
    const unsigned grid = (unsigned)(g.num_tiles < num_sms() ? g.num_tiles : num_sms());
    launch_tc(kern, grid, num_threads<CostVolProducer>(), smem, to_stream(stream), g, pa, ea);
    KDPC_RETURN_LAST();
}
