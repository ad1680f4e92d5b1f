// Weight gradient of every Linear / 1x1 convolution of the training path on tcgen05 (sm_100a):
//
//        dW[n, k] = sum_m dY[m, n] * X[m, k]          (what autograd computes for nn.Linear / nn.Conv1d(k=1):
//                                                       reference pointconv_util.py:20-54, 223, 250 via loss.backward(),
//                                                       distilTrain.py:180)
//
// The reduction runs over the ROWS m (65 536 .. 262 144 points), the output is small (N <= 256 by K <= 2096), so both
// operands are "MN-major" for the tensor core: dY[m, :] and X[m, :] are contiguous along the NON-reduced index.  No
// transposition anywhere: producers read 8 consecutive floats of a row, split them into bf16 hi / lo and store ONE
// 16-byte unit into the canonical MN-major SWIZZLE_128B tile (atom = 64 MN elements x 8 reduction rows, 16-byte units
// XOR-swizzled by the row, atoms LBO apart along MN and SBO apart along the reduction), and tcgen05.mma is issued with
// a_major = b_major = MN.  fp32 accuracy from three bf16 MMAs per step (hi*hi + hi*lo + lo*hi), fp32 accumulation in TMEM.
// (torch's fp32 mm for this product runs on the CUDA cores - cutlass_80_simt_sgemm - and was 28 % of the KD step.)
//
// Work decomposition: output tiles of 128 (n) x KT <= 256 (k); the row range is split over CTAs so that ~all SMs work
// (few output tiles, very long reduction); every CTA leaves its partial tile in a workspace and dw_reduce_kernel adds
// the partials in split order: deterministic.
#include "tc_common.cuh"

namespace kdpc {
namespace tc {

constexpr int DW_MCHUNK = 64;                       // reduction rows per pipeline stage (4 UMMA K-steps)
constexpr int DW_TN = 128;                          // n rows of an output tile (UMMA M)
constexpr int DW_KT = 256;                          // k columns of an output tile (UMMA N), multiple of 64 in smem
constexpr int DW_SBO = 1024;                        // bytes between 8-row reduction groups (one swizzle atom)
constexpr int DW_LBO = DW_MCHUNK / 8 * DW_SBO;      // bytes between 64-element MN groups: 8 atoms = 8192
constexpr int DW_A_PART = DW_TN / 64 * DW_LBO;      // one bf16 part (hi or lo) of the dY tile: 16 KB
constexpr int DW_PRODUCERS = 256;
constexpr int DW_THREADS = DW_PRODUCERS + 32 + 128; // producers + MMA warp + 4 epilogue warps

// Shared-memory matrix descriptor, MN-major SWIZZLE_128B (cute UMMA::make_umma_desc<Major::MN>):
//   [0,14) addr >> 4 | [16,30) LBO >> 4 (between 64-element MN groups) | [32,46) SBO >> 4 (between 8-row K groups)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor: bf16 x bf16 -> fp32, BOTH operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int m, int n) {
    return make_idesc_bf16(m, n) | (1u << 15) | (1u << 16);
}
// byte offset of the 16-byte unit holding MN elements [8c, 8c+8) of reduction row r inside an MN-major tile
__device__ __forceinline__ uint32_t mn_offset(int mn_unit, int r) {
    const int g = mn_unit >> 3, c = mn_unit & 7;                 // 64-element group, unit inside it
    return (uint32_t)(g * DW_LBO + (r >> 3) * DW_SBO + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

struct DwArgs {
    const float *dy;      // [M, ldy]
    const float *x;       // [M, ldx]
    long long m;
    int n, k, ldy, ldx;
    int k_eff;                       // k + 1 when the bias gradient rides along as an extra all-ones column of X
    int n_tiles, k_tiles, splits;
    long long rows_per_split;        // multiple of DW_MCHUNK
    float *partial;                  // [splits][n_tiles * 128][k_tiles * 256]
    int raw_stages;                  // > 0: narrow contiguous operands - row chunks staged by 1-D bulk TMA, raw_stages - 1 chunks ahead
};
constexpr int DW_MAX_RAW = 4;

// unit = 8 consecutive floats of one source row -> (hi, lo) 16-byte units; `width` valid floats from `col0`
__device__ __forceinline__ void dw_load_unit(const float *rowp, bool row_ok, int col0, int width, bool vec, float (&v)[8]) {
    if (row_ok && vec && col0 + 8 <= width) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(rowp + col0));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(rowp + col0 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (row_ok && col0 + j < width) ? __ldg(rowp + col0 + j) : 0.f;
    }
}

__global__ void __launch_bounds__(DW_THREADS, 1)
dw_tc_kernel(const DwArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[2], empty[2], done_bar, raw_full[DW_MAX_RAW];
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

    // work item -> (n tile, k tile, split)
    const int tiles = a.n_tiles * a.k_tiles;
    const int split = blockIdx.x / tiles;
    const int tile = blockIdx.x - split * tiles;
    const int nt = tile / a.k_tiles, kt = tile - nt * a.k_tiles;
    const int n0 = nt * DW_TN, k0 = kt * DW_KT;
    const int kw = min(DW_KT, a.k_eff - k0);                     // valid k columns of this tile (incl. the ones column)
    const int kgroups = (kw + 63) >> 6;                          // 64-column groups staged in shared memory
    const int umma_n = (kw + 15) & ~15;                          // UMMA N (multiple of 16)
    const int b_part = kgroups * DW_LBO;                         // one bf16 part of the X tile
    const int stage_bytes = 2 * DW_A_PART + 2 * b_part;
    const long long m_begin = (long long)split * a.rows_per_split;
    const long long m_end = min(a.m, m_begin + a.rows_per_split);
    const int chunks = m_end > m_begin ? (int)((m_end - m_begin + DW_MCHUNK - 1) / DW_MCHUNK) : 0;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&full[s], DW_PRODUCERS / 32); mbar_init(&empty[s], 1); }
        mbar_init(&done_bar, 1);
        for (int r = 0; r < DW_MAX_RAW; ++r) mbar_init(&raw_full[r], 1);     // expect_tx arrival of producer thread 0
        mbar_fence_init();
    }
    if (warp == DW_PRODUCERS / 32) tmem_alloc(&tmem_base_smem, 256u);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < DW_PRODUCERS / 32) {
        // ================= producers: fp32 rows -> bf16 hi/lo, MN-major SW128 tiles =================
        const bool vec_y = (a.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dy) & 15) == 0 && (n0 & 3) == 0;
        const bool vec_x = (a.ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
        const int b_upr = kgroups * 8;                           // units per row of the X tile
        if (a.raw_stages > 0) {
            // Narrow layers over very many rows (cost-volume / positional layers: N, K <= 64 over 0.5 - 2 M rows).  The
            // register-staged loop below has ONE 64-row chunk (<= 17 KB) per memory round trip per CTA and ran 8x above
            // the HBM bound.  Here the operands are contiguous ([M, n] and [M, k], M % 64 == 0, one output tile): a chunk
            // of each is ONE 1-D bulk TMA copy (64 n * 4 and 64 k * 4 bytes) issued raw_stages - 1 chunks ahead by
            // thread 0, and the producers convert from shared memory only.  (2 M x 32 x 32: 666 -> 341 us.  Knock-outs: one
            // MMA of three: no change; no conversion: 185 us for either narrow shape - ~1650 cycles of barrier / wait /
            // fence / commit per 64-row chunk are the floor of this pipeline, the conversion adds ~1400.)
            const int R = a.raw_stages;
            const uint32_t ybytes = (uint32_t)DW_MCHUNK * a.n * 4u, xbytes = (uint32_t)DW_MCHUNK * a.k * 4u;
            const uint32_t raw_bytes = (ybytes + xbytes + 15u) & ~15u;
            unsigned char *raw_base = smem + 2 * (size_t)stage_bytes;
            auto issue = [&](int c) {
                const int slot = c % R;
                unsigned char *dst = raw_base + (size_t)slot * raw_bytes;
                const long long row0 = m_begin + (long long)c * DW_MCHUNK;
                mbar_expect_tx(&raw_full[slot], ybytes + xbytes);
                tma_load_1d(dst, a.dy + row0 * a.n, ybytes, &raw_full[slot]);
                tma_load_1d(dst + ybytes, a.x + row0 * a.k, xbytes, &raw_full[slot]);
            };
            if (tid == 0)
                for (int c = 0; c < R - 1 && c < chunks; ++c) issue(c);
            const int a_live = (a.n + 7) >> 3, b_live = min(b_upr, (kw + 7) >> 3);
            const bool vec_n = (a.n & 3) == 0, vec_k = (a.k & 3) == 0;
            for (int c = 0; c < chunks; ++c) {
                const int s = c & 1;
                // every producer is done converting chunk c - 1: its raw slot may be refilled
                asm volatile("bar.sync 1, %0;" ::"n"(DW_PRODUCERS) : "memory");
                if (tid == 0 && c + R - 1 < chunks) issue(c + R - 1);
                mbar_wait(&empty[s], ((c >> 1) & 1) ^ 1);
                unsigned char *st = smem + (size_t)s * stage_bytes;
                unsigned char *a_hi = st, *a_lo = st + DW_A_PART, *b_hi = st + 2 * DW_A_PART, *b_lo = b_hi + b_part;
                const bool first_fill = c < 2;
                if (first_fill) {                                 // all-zero padding units, once per operand stage
                    const uint4 z = make_uint4(0, 0, 0, 0);
                    for (int i = tid; i < DW_MCHUNK * (DW_TN / 8); i += DW_PRODUCERS) {
                        const int r = i / (DW_TN / 8), u = i - r * (DW_TN / 8);
                        if (u >= a_live) { const uint32_t off = mn_offset(u, r); *reinterpret_cast<uint4 *>(a_hi + off) = z; *reinterpret_cast<uint4 *>(a_lo + off) = z; }
                    }
                    for (int i = tid; i < DW_MCHUNK * b_upr; i += DW_PRODUCERS) {
                        const int r = i / b_upr, u = i - r * b_upr;
                        if (u >= b_live) { const uint32_t off = mn_offset(u, r); *reinterpret_cast<uint4 *>(b_hi + off) = z; *reinterpret_cast<uint4 *>(b_lo + off) = z; }
                    }
                }
                mbar_wait(&raw_full[c % R], (uint32_t)((c / R) & 1));
                const float *ry = reinterpret_cast<const float *>(raw_base + (size_t)(c % R) * raw_bytes);
                const float *rx = reinterpret_cast<const float *>(raw_base + (size_t)(c % R) * raw_bytes + ybytes);
                // unit = 8 consecutive floats of one row; consecutive threads take consecutive units (conflict-free reads)
                auto load_unit = [&](const float *row, int col0, int width, bool vec, float (&v)[8]) {
                    if (vec && col0 + 8 <= width) {
                        const float4 p = *reinterpret_cast<const float4 *>(row + col0), q = *reinterpret_cast<const float4 *>(row + col0 + 4);
                        v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = col0 + j < width ? row[col0 + j] : 0.f;
                    }
                };
                for (int i = tid; i < DW_MCHUNK * a_live; i += DW_PRODUCERS) {
                    const int r = i / a_live, u = i - r * a_live;
                    float v[8];
                    load_unit(ry + r * a.n, u * 8, a.n, vec_n, v);
                    uint4 hi, lo;
                    split8(v, hi, lo);
                    const uint32_t off = mn_offset(u, r);
                    *reinterpret_cast<uint4 *>(a_hi + off) = hi;
                    *reinterpret_cast<uint4 *>(a_lo + off) = lo;
                }
                for (int i = tid; i < DW_MCHUNK * b_live; i += DW_PRODUCERS) {
                    const int r = i / b_live, u = i - r * b_live;
                    float v[8];
                    load_unit(rx + r * a.k, u * 8, a.k, vec_k, v);
                    const int oc = a.k - u * 8;                   // the all-ones column (bias gradient) inside this unit?
                    if (a.k_eff != a.k && oc >= 0 && oc < 8) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (q == oc) v[q] = 1.f;
                    }
                    uint4 hi, lo;
                    split8(v, hi, lo);
                    const uint32_t off = mn_offset(u, r);
                    *reinterpret_cast<uint4 *>(b_hi + off) = hi;
                    *reinterpret_cast<uint4 *>(b_lo + off) = lo;
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        } else
        for (int c = 0; c < chunks; ++c) {
            const int s = c & 1;
            mbar_wait(&empty[s], ((c >> 1) & 1) ^ 1);
            unsigned char *st = smem + (size_t)s * stage_bytes;
            unsigned char *a_hi = st, *a_lo = st + DW_A_PART, *b_hi = st + 2 * DW_A_PART, *b_lo = b_hi + b_part;
            const long long row0 = m_begin + (long long)c * DW_MCHUNK;
            // Units that lie entirely beyond the tile's valid columns (n >= N, k >= K_eff) are all-zero in EVERY chunk:
            // they are written when a stage is filled for the first time (c < 2) and skipped afterwards (for a 32-wide
            // layer three quarters of the dY tile are such padding).
            const bool first_fill = c < 2;
            const int a_live = min(DW_TN / 8, (a.n - n0 + 7) >> 3);   // live units per dY row
            const int b_live = min(b_upr, (kw + 7) >> 3);              // live units per X row
            // dY tile: 64 rows x 128 columns (n0 ..)
            {
                const int upr = first_fill ? DW_TN / 8 : a_live;
                const int units = DW_MCHUNK * upr;
                for (int i0 = tid; i0 < units; i0 += 4 * DW_PRODUCERS) {
                    float v[4][8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = i0 + j * DW_PRODUCERS;
                        const int r = i / upr, u = i - r * upr;
                        const long long row = row0 + r;
                        dw_load_unit(a.dy + row * a.ldy + n0, i < units && row < m_end, u * 8, a.n - n0, vec_y, v[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = i0 + j * DW_PRODUCERS;
                        if (i < units) {
                            const int r = i / upr, u = i - r * upr;
                            uint4 hi, lo;
                            split8(v[j], hi, lo);
                            const uint32_t off = mn_offset(u, r);
                            *reinterpret_cast<uint4 *>(a_hi + off) = hi;
                            *reinterpret_cast<uint4 *>(a_lo + off) = lo;
                        }
                    }
                }
            }
            // X tile: 64 rows x (kgroups * 64) columns (k0 ..)
            {
                const int upr = first_fill ? b_upr : b_live;
                const int units = DW_MCHUNK * upr;
                for (int i0 = tid; i0 < units; i0 += 4 * DW_PRODUCERS) {
                    float v[4][8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = i0 + j * DW_PRODUCERS;
                        const int r = i / upr, u = i - r * upr;
                        const long long row = row0 + r;
                        dw_load_unit(a.x + row * a.ldx + k0, i < units && row < m_end, u * 8, a.k - k0, vec_x && (k0 & 3) == 0, v[j]);
                        // db = sum_m dY[m, :] rides along as column k of X == 1 (valid rows only)
                        const int oc = a.k - k0 - u * 8;              // position of the ones column inside this unit
                        if (a.k_eff != a.k && oc >= 0 && oc < 8 && i < units && row < m_end) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                if (q == oc) v[j][q] = 1.f;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = i0 + j * DW_PRODUCERS;
                        if (i < units) {
                            const int r = i / upr, u = i - r * upr;
                            uint4 hi, lo;
                            split8(v[j], hi, lo);
                            const uint32_t off = mn_offset(u, r);
                            *reinterpret_cast<uint4 *>(b_hi + off) = hi;
                            *reinterpret_cast<uint4 *>(b_lo + off) = lo;
                        }
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }
    } else if (warp == DW_PRODUCERS / 32) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_bf16_mn(DW_TN, umma_n);
        const uint32_t base = smem_u32(smem);
        for (int c = 0; c < chunks; ++c) {
            const int s = c & 1;
            mbar_wait(&full[s], (c >> 1) & 1);
            fence_after_sync();
            const uint32_t a_hi = base + (uint32_t)s * (uint32_t)stage_bytes, a_lo = a_hi + DW_A_PART;
            const uint32_t b_hi = a_hi + 2 * DW_A_PART, b_lo = b_hi + (uint32_t)b_part;
            const bool leader = elect_one();
#pragma unroll
            for (int kk = 0; kk < DW_MCHUNK / UMMA_K; ++kk) {
                const uint32_t off = (uint32_t)kk * 2u * DW_SBO;              // 16 reduction rows = two 8-row groups
                const uint64_t dah = make_smem_desc_mn_sw128(a_hi + off, DW_LBO, DW_SBO), dal = make_smem_desc_mn_sw128(a_lo + off, DW_LBO, DW_SBO);
                const uint64_t dbh = make_smem_desc_mn_sw128(b_hi + off, DW_LBO, DW_SBO), dbl = make_smem_desc_mn_sw128(b_lo + off, DW_LBO, DW_SBO);
                if (leader) {
                    umma_bf16(tmem_base, dah, dbh, idesc, (c != 0) || (kk != 0));
                    umma_bf16(tmem_base, dah, dbl, idesc, 1);
                    umma_bf16(tmem_base, dal, dbh, idesc, 1);
                }
            }
            if (leader) {
                umma_commit(&empty[s]);
                if (c == chunks - 1) umma_commit(&done_bar);
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue: partial tile -> workspace =================
        const int quarter = warp & 3;
        float *prow = a.partial + ((size_t)split * a.n_tiles * DW_TN + (size_t)n0 + quarter * 32 + lane) * ((size_t)a.k_tiles * DW_KT) + k0;
        if (chunks > 0) {
            mbar_wait(&done_bar, 0);
            fence_after_sync();
        }
        for (int c0 = 0; c0 < umma_n; c0 += 32) {
            float v[32];
            if (chunks > 0) {
                tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (c0 + j < umma_n)
                    *reinterpret_cast<float4 *>(prow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        fence_before_sync();
    }
    __syncthreads();
    if (warp == DW_PRODUCERS / 32) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 256u);
    }
}

// dW for layers with a tiny input width (the 3 -> D positional encodings over B*N*K rows: k + bias <= 4): the product is
// k + 1 weighted column sums of dY - a tensor-core tile would be 98 % padding and the MN-major pipeline above runs such
// shapes at ~1 TB/s.  CUDA cores: thread = (4 output columns, row lane), float4 loads of dY (coalesced), X broadcast per
// row, fp32 accumulators walked in row order; the row lanes of a CTA are summed in lane order through shared memory and
// every CTA leaves its [n x KE] partial in the SAME workspace layout as dw_tc_kernel, so dw_reduce_kernel finishes both.
constexpr int DWS_THREADS = 512;
template <int KE>                                          // k + (bias ? 1 : 0) <= KE = 4
__global__ void __launch_bounds__(DWS_THREADS)
dw_small_k_kernel(const DwArgs a) {
    extern __shared__ float4 sred[];                       // [row lanes][column groups][KE]
    const int cg = a.n >> 2;                               // column groups of 4 (n % 4 == 0)
    const int lanes = DWS_THREADS / cg;                    // row lanes
    const int c = threadIdx.x % cg, rl = threadIdx.x / cg;
    const long long m_begin = (long long)blockIdx.x * a.rows_per_split;
    const long long m_end = min(a.m, m_begin + a.rows_per_split);
    float4 acc[KE];
#pragma unroll
    for (int j = 0; j < KE; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool bias = a.k_eff != a.k;
    if (rl < lanes) {
        long long row = m_begin + rl;
        for (; row + 3LL * lanes < m_end; row += 4LL * lanes) {        // four rows in flight per thread
            float4 d[4];
            float xv[4][KE];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long r = row + (long long)u * lanes;
                d[u] = ld_stream_f4(reinterpret_cast<const float4 *>(a.dy + r * a.ldy) + c);
#pragma unroll
                for (int j = 0; j < KE; ++j) xv[u][j] = j < a.k ? __ldg(a.x + r * a.ldx + j) : 1.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < KE; ++j) {
                    acc[j].x = fmaf(d[u].x, xv[u][j], acc[j].x); acc[j].y = fmaf(d[u].y, xv[u][j], acc[j].y);
                    acc[j].z = fmaf(d[u].z, xv[u][j], acc[j].z); acc[j].w = fmaf(d[u].w, xv[u][j], acc[j].w);
                }
        }
        for (; row < m_end; row += lanes) {
            const float4 d = ld_stream_f4(reinterpret_cast<const float4 *>(a.dy + row * a.ldy) + c);
#pragma unroll
            for (int j = 0; j < KE; ++j) {
                const float xv = j < a.k ? __ldg(a.x + row * a.ldx + j) : 1.f;
                acc[j].x = fmaf(d.x, xv, acc[j].x); acc[j].y = fmaf(d.y, xv, acc[j].y);
                acc[j].z = fmaf(d.z, xv, acc[j].z); acc[j].w = fmaf(d.w, xv, acc[j].w);
            }
        }
#pragma unroll
        for (int j = 0; j < KE; ++j) sred[(rl * cg + c) * KE + j] = acc[j];
    }
    __syncthreads();
    // thread (column group, j): its column group's lanes in lane order
    if (threadIdx.x < cg * KE) {
        const int cc = threadIdx.x / KE, j = threadIdx.x - cc * KE;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < lanes; ++l) {
            const float4 v = sred[(l * cg + cc) * KE + j];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        if (j < a.k_eff) {
            float *p = a.partial + ((size_t)blockIdx.x * a.n_tiles * DW_TN + (size_t)cc * 4) * ((size_t)a.k_tiles * DW_KT) + j;
            const size_t ld = (size_t)a.k_tiles * DW_KT;
            p[0] = t.x; p[ld] = t.y; p[2 * ld] = t.z; p[3 * ld] = t.w;
        }
    }
    (void)bias;
}

// dW[n, k] = sum over the splits' partial tiles; column k (when present) is the bias gradient db[n].  One WARP per group of
// four output columns: lane l adds splits l, l + 32, .. in order, then a fixed butterfly adds the 32 lanes (deterministic;
// one thread per group walking up to 148 splits one after the other was 23 us per call, 1.8 ms of the KD step).
__global__ void __launch_bounds__(256)
dw_reduce_kernel(int n, int k, int k_eff, int splits, long long split_stride, int ldp, const float *__restrict__ partial,
                 float *__restrict__ dw, int lddw, float *__restrict__ db) {
    const int kq = (k_eff + 3) >> 2;
    const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= (long long)n * kq) return;                   // (whole warps)
    const int row = (int)(t / kq), c0 = (int)(t - (long long)row * kq) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *p = partial + (size_t)row * ldp + c0;
    for (int s = lane; s < splits; s += 32) {
        const float4 v = *reinterpret_cast<const float4 *>(p + (size_t)s * split_stride);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (lane != 0) return;
    const float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (c0 + j < k) dw[(size_t)row * lddw + c0 + j] = o[j];
        else if (c0 + j == k && k_eff != k) db[row] = o[j];
    }
}

static inline void dw_plan(long long m, int n, int k, bool bias, DwArgs &a) {
    a.k_eff = k + (bias ? 1 : 0);
    a.n_tiles = (n + DW_TN - 1) / DW_TN;
    a.k_tiles = (a.k_eff + DW_KT - 1) / DW_KT;
    const int tiles = a.n_tiles * a.k_tiles;
    const long long mchunks = (m + DW_MCHUNK - 1) / DW_MCHUNK;
    // one wave of CTAs: every split leaves a [128 x 256] partial tile that dw_reduce_kernel has to read again, so more
    // splits than SMs only add reduction traffic (2 waves: 296 partial tiles = 19 MB read back for a 64 KB gradient)
    long long want = (device_sms() + tiles - 1) / tiles;
    if (want > (mchunks + 3) / 4) want = (mchunks + 3) / 4;          // >= 4 chunks of 64 rows per CTA
    if (want > mchunks) want = mchunks;
    if (want < 1) want = 1;
    const long long cps = (mchunks + want - 1) / want;               // chunks per split
    a.rows_per_split = cps * DW_MCHUNK;
    a.splits = (int)((mchunks + cps - 1) / cps);
}

}  // namespace tc
}  // namespace kdpc

using namespace kdpc;
using namespace kdpc::tc;

static int kdpc_dw_async = 1;
/* A/B switch for measurements: 0 = register-staged row fetch for every shape (same results) */
KDPC_API void kdpc_linear_dw_set_async(int on) { kdpc_dw_async = on; }

KDPC_API long long kdpc_linear_dw_ws_bytes(long long m, int n, int k) {
    if (m <= 0 || n <= 0 || k <= 0) return 0;
    DwArgs a{};
    dw_plan(m, n, k, true, a);                               // (sized for the variant with the bias column)
    return (long long)a.splits * a.n_tiles * DW_TN * a.k_tiles * DW_KT * (long long)sizeof(float);
}

/* dW[n,k] = sum_m dY[m,n] X[m,k]  (weight gradient of y = x W^T) and, when db != NULL, db[n] = sum_m dY[m,n] from the same
 * launch.  dy [M,ldy], x [M,ldx], dw [N,lddw]; ws: kdpc_linear_dw_ws_bytes. */
KDPC_API int kdpc_linear_dw(long long m, int n, int k, const float *dy, int ldy, const float *x, int ldx, void *ws,
                            float *dw, int lddw, float *db, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(dy && x && ws && dw && m > 0 && n > 0 && k > 0 && ldy >= n && ldx >= k && lddw >= k);
    if ((reinterpret_cast<uintptr_t>(ws) % 16) != 0) return KDPC_EINVAL;
    DwArgs a{};
    a.dy = dy; a.x = x; a.m = m; a.n = n; a.k = k; a.ldy = ldy; a.ldx = ldx;
    dw_plan(m, n, k, db != nullptr, a);
    a.partial = reinterpret_cast<float *>(ws);
    size_t smem = 2 * (2 * (size_t)DW_A_PART + 2 * (size_t)(DW_KT / 64) * DW_LBO) + 1024;
    if (smem < 226 * 1024) smem = 226 * 1024;              // (the TMA-staged path puts its raw ring behind the operand stages)
    a.raw_stages = 0;
    cudaStream_t st = to_stream(stream);
    const int ldp = a.k_tiles * DW_KT;
    const long long split_stride = (long long)a.n_tiles * DW_TN * ldp;
    const long long total = (long long)n * ((a.k_eff + 3) / 4);
    if (kdpc_dw_async && a.k_eff <= 4 && n <= 256 && (n & 3) == 0 && (ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
        // tiny input width: weighted column sums on the CUDA cores (same workspace layout, same reduce kernel)
        const int lanes = DWS_THREADS / (n / 4);
        dw_small_k_kernel<4><<<(unsigned)a.splits, DWS_THREADS, (size_t)lanes * (n / 4) * 4 * sizeof(float4), st>>>(a);
        dw_reduce_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, st>>>(n, k, a.k_eff, a.splits, split_stride, ldp, a.partial, dw, lddw, db);
        KDPC_RETURN_LAST();
    }
    const int kgroups = (a.k_eff + 63) / 64;
    if (kdpc_dw_async && a.n_tiles == 1 && a.k_tiles == 1 && ldy == n && ldx == k && m % DW_MCHUNK == 0 && kgroups <= 2 &&
        ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x)) & 15) == 0) {
        // narrow contiguous operands: raw row chunks by bulk TMA next to the two operand stages (<= two 64-column X groups)
        const size_t stages = 2 * (2 * (size_t)DW_A_PART + 2 * (size_t)kgroups * DW_LBO);
        const size_t raw = ((size_t)DW_MCHUNK * (n + k) * 4 + 15) & ~(size_t)15;
        const long long fit = (long long)((smem - 1024 - stages) / raw);
        if (fit >= 2) a.raw_stages = (int)(fit < DW_MAX_RAW ? fit : DW_MAX_RAW);
    }
    KDPC_ENSURE_SMEM(dw_tc_kernel, (int)smem);
    const unsigned grid = (unsigned)(a.n_tiles * a.k_tiles * a.splits);
    dw_tc_kernel<<<grid, DW_THREADS, smem, st>>>(a);
    dw_reduce_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, st>>>(n, k, a.k_eff, a.splits, split_stride, ldp, a.partial, dw, lddw, db);
    KDPC_RETURN_LAST();
}
