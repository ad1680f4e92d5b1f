// Fused linear layers on tcgen05 (sm_100a):  y = clamp(act((x W^T) * scale + shift)) (+ residual)
//
// Replaces the cuBLAS fp32 SIMT GEMMs behind the reference's 1x1 convolutions / nn.Linear
// (Conv1d/Conv2d wrappers pointconv_util.py:20-54, PointConv.linear :250, cross_t* :1800-1811,
// SceneFlowEstimatorResidual.fc :2234) with ONE kernel per layer: bias, eval-mode BatchNorm,
// LeakyReLU, clamp and the residual add ride in the epilogue.  Accuracy: bf16 hi/lo split, three
// MMAs per K-step, fp32 accumulation in TMEM (tc_common.cuh).
//
// Tiny layers (K < 16 or N < 16: level0 3->32, fc 64->3) are issue-/HBM-bound, not GEMM-shaped:
// they use a SIMT kernel with the same epilogue.
#include "tc_gemm.cuh"

namespace kdpc {
namespace tc {

// Pack fp32 weights [N, K_src] (row-major, nn.Linear / 1x1 conv layout) into per-chunk SWIZZLE_128B
// bf16 hi/lo tile images: out[chunk][part][n_pad][128 B].
//   mode 0: packed column kc <- source column kc.
//   mode 1 (PointConv, weightnet width wn): the fused kernel orders channels as
//           [dx, dy, dz, zero, features 0..D-1] (so that feature rows gather as aligned float4 and the
//           coordinate chunk comes first), the reference concatenates [dx,dy,dz, features]
//           (pointconv_util.py:153,178) and flattens c-major with wn innermost (:249).
//           packed kc = c'*wn + w  <-  source (c*wn + w) with c = c' for c' < 3, zero for c' = 3, c = c'-1 above.
__global__ void pack_weight_kernel(int n, int k_src, int n_pad, int num_chunks, int mode, int d, int wn,
                                   const float *__restrict__ w, unsigned char *__restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // (chunk, row, unit)
    const long long total = (long long)num_chunks * n_pad * 8;
    if (e >= total) return;
    const int u = (int)(e & 7);
    const int row = (int)((e >> 3) % n_pad);
    const int chunk = (int)((e >> 3) / n_pad);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int kc = chunk * CHUNK_K + u * 8 + j;
        int src = -1;
        if (mode == 0) {
            src = kc < k_src ? kc : -1;
        } else {
            const int cp = kc / wn, wi = kc - cp * wn;
            if (cp < 3) src = cp * wn + wi;
            else if (cp > 3 && cp < d + 4) src = (cp - 1) * wn + wi;
        }
        v[j] = (row < n && src >= 0) ? w[(size_t)row * k_src + src] : 0.f;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    unsigned char *base = out + (size_t)chunk * (2 * n_pad * 128);
    const uint32_t off = sw128_offset(row, u);
    *reinterpret_cast<uint4 *>(base + off) = hi;
    *reinterpret_cast<uint4 *>(base + (size_t)n_pad * 128 + off) = lo;
}

// SIMT path for tiny layers: thread = (row, group of 4 outputs).
__global__ void __launch_bounds__(256)
linear_simt_kernel(long long m, int n, int k, const float *__restrict__ x, int ldx, const float *__restrict__ w,
                   const float *__restrict__ scale, const float *__restrict__ shift, float slope, float lo, float hi,
                   const float *__restrict__ residual, float *__restrict__ out, int ldo) {
    const int ng = (n + 3) >> 2;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m * ng) return;
    const long long row = e / ng;
    const int n0 = (int)(e - row * ng) * 4;
    const float *xr = x + row * ldx;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int kk = 0; kk < k; ++kk) {
        const float xv = __ldg(xr + kk);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n0 + j < n) acc[j] = fmaf(xv, __ldg(w + (size_t)(n0 + j) * k + kk), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int col = n0 + j;
        if (col < n) {
            float y = acc[j];
            if (scale) y *= __ldg(scale + col);
            if (shift) y += __ldg(shift + col);
            y = y > 0.f ? y : y * slope;
            if (lo <= hi) y = fminf(fmaxf(y, lo), hi);
            if (residual) y += __ldg(residual + row * ldo + col);
            out[row * ldo + col] = y;
        }
    }
}


// Two shapes of the positional-encoding layers over B*N*K rows, where the generic kernel above runs at 1/4 - 1/7 of the
// HBM bound (one thread per (row, 4 outputs) re-reads its weights for every row and walks its x row scalar by scalar):
//   * k <= 4 (3 -> D forward): thread = (row, 4 outputs) as above, but a thread keeps its 4 x k weights and its affine in
//     registers and walks rows grid-stride: the loop is 1-3 loads, 4k FMAs and one 16-byte store per row;
template <int KMAX>
__global__ void __launch_bounds__(256)
linear_smallk_kernel(long long m, int n, int k, const float *__restrict__ x, int ldx, const float *__restrict__ w,
                     const float *__restrict__ scale, const float *__restrict__ shift, float slope, float lo, float hi,
                     float *__restrict__ out, int ldo) {
    const int ng = n >> 2;                                    // n % 4 == 0
    const int g = threadIdx.x % ng, rl = threadIdx.x / ng, lanes = blockDim.x / ng;
    if (rl >= lanes) return;
    float wr[4][KMAX], sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk) wr[j][kk] = kk < k ? __ldg(w + (size_t)(g * 4 + j) * k + kk) : 0.f;
        sc[j] = scale ? __ldg(scale + g * 4 + j) : 1.f;
        sh[j] = shift ? __ldg(shift + g * 4 + j) : 0.f;
    }
    const bool clampd = lo <= hi;
    for (long long row = (long long)blockIdx.x * lanes + rl; row < m; row += (long long)gridDim.x * lanes) {
        float xv[KMAX];
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk) xv[kk] = kk < k ? __ldg(x + row * ldx + kk) : 0.f;
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int kk = 0; kk < KMAX; ++kk) acc = fmaf(xv[kk], wr[j][kk], acc);     // same order as the generic kernel
            float t = acc * sc[j];
            t += sh[j];
            t = t > 0.f ? t : t * slope;
            if (clampd) t = fminf(fmaxf(t, lo), hi);
            y[j] = t;
        }
        st_stream_f4(reinterpret_cast<float4 *>(out + row * ldo) + g, make_float4(y[0], y[1], y[2], y[3]));
    }
}
//   * n <= 4 (D -> 3: flow heads, and the input gradient of the 3 -> D layers): 8 lanes share a row, each loads float4s of
//     it (coalesced 128-byte requests), keeps partial dot products for the n outputs and the 8 lanes are summed by a
//     fixed shuffle tree; lane 0 of the group applies the epilogue and stores.
template <int NMAX>
__global__ void __launch_bounds__(256)
linear_smalln_kernel(long long m, int n, int k, const float *__restrict__ x, int ldx, const float *__restrict__ w,
                     const float *__restrict__ scale, const float *__restrict__ shift, float slope, float lo, float hi,
                     const float *__restrict__ residual, float *__restrict__ out, int ldo) {
    const int sub = threadIdx.x & 7;
    const long long rows_per_cta = blockDim.x >> 3;
    const bool clampd = lo <= hi;
    for (long long row = (long long)blockIdx.x * rows_per_cta + (threadIdx.x >> 3);; row += (long long)gridDim.x * rows_per_cta) {
        // (whole warps leave together: rows_per_cta * gridDim.x strides keep a warp's 4 rows consecutive)
        const bool live = row < m;
        if (__all_sync(0xffffffffu, !live)) break;
        float acc[NMAX];
#pragma unroll
        for (int j = 0; j < NMAX; ++j) acc[j] = 0.f;
        if (live) {
            const float4 *xr = reinterpret_cast<const float4 *>(x + row * ldx);
            for (int q = sub; q < (k >> 2); q += 8) {         // k % 4 == 0
                const float4 v = ld_stream_f4(xr + q);
#pragma unroll
                for (int j = 0; j < NMAX; ++j) {
                    if (j < n) {
                        const float4 ww = __ldg(reinterpret_cast<const float4 *>(w + (size_t)j * k) + q);
                        acc[j] = fmaf(v.x, ww.x, acc[j]); acc[j] = fmaf(v.y, ww.y, acc[j]);
                        acc[j] = fmaf(v.z, ww.z, acc[j]); acc[j] = fmaf(v.w, ww.w, acc[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
        }
        if (live && sub == 0) {
#pragma unroll
            for (int j = 0; j < NMAX; ++j) {
                if (j < n) {
                    float y = acc[j];
                    if (scale) y *= __ldg(scale + j);
                    if (shift) y += __ldg(shift + j);
                    y = y > 0.f ? y : y * slope;
                    if (clampd) y = fminf(fmaxf(y, lo), hi);
                    if (residual) y += __ldg(residual + row * ldo + j);
                    out[row * ldo + j] = y;
                }
            }
        }
    }
}

}  // namespace tc
}  // namespace kdpc

using namespace kdpc;
using namespace kdpc::tc;

// cuTensorMapEncodeTiled through the runtime (no libcuda link dependency); nullptr when the driver lacks it
typedef CUresult (*kdpc_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                         const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static kdpc_encode_tiled_fn tensor_map_encoder() {
    static kdpc_encode_tiled_fn fn = nullptr;
    static int tried = 0;
    if (!tried) {
        tried = 1;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<kdpc_encode_tiled_fn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}
// [M, K] fp32 rows of pitch ldx: box = 32 columns x 128 rows, SWIZZLE_128B, out-of-range elements read as zero
static bool make_row_tensor_map(CUtensorMap *tm, const float *x, long long m, int k, int ldx) {
    kdpc_encode_tiled_fn enc = tensor_map_encoder();
    if (enc == nullptr) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)m};
    const cuuint64_t gstride[1] = {(cuuint64_t)ldx * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)PlainTmaProducer::BOX_COLS, (cuuint32_t)TILE_M};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_tc_pdl = 0;              // measured neutral (single graph 4.897 vs 4.896 ms, pipeline 2246 vs 2263 pairs/s): off
KDPC_API int kdpc_tc_pdl_enabled(void) { return g_tc_pdl; }
/* A/B switch for measurements: 1 = programmatic dependent launch of the tcgen05 kernels (same results) */
KDPC_API void kdpc_tc_set_pdl(int on) { g_tc_pdl = on; }
static int kdpc_linear_split_n = 1;
static int g_tc_async = 2;              // 0 = synchronous producers, 1 = cp.async (LDGSTS) rows, 2 = tensor-map TMA rows (default)
KDPC_API int kdpc_tc_async_enabled(void) { return g_tc_async; }
KDPC_API void kdpc_tc_set_async(int on) { g_tc_async = on; }
/* A/B switch for measurements: 0 = small-M layers always by split-K + reduce (results differ in the last bits: other summation order) */
KDPC_API void kdpc_linear_set_split_n(int on) { kdpc_linear_split_n = on; }
static void *g_tc_trace = nullptr;
KDPC_API void *kdpc_tc_trace_buffer(void) { return g_tc_trace; }
/* debug (tools/trace_*.py): every tcgen05 kernel launched while a buffer is set writes CTA 0's per-iteration clock64
 * stamps into it (200 x 16 int64, device memory); NULL = off */
KDPC_API void kdpc_tc_set_trace(void *p) { g_tc_trace = p; }

// rows of a packed weight: multiples of 16 up to 256 outputs; WIDE layers (n > 256: level3_1, the input gradients of the
// PointConv linears) are padded to whole 128-row column blocks (kdpc_linear_tc walks them as split-N work items)
static inline int packed_rows(int n) { return n > 256 ? (n + 127) / 128 * 128 : (n + 15) / 16 * 16; }

KDPC_API long long kdpc_packed_weight_bytes(int n, int k_packed) {
    const int n_pad = packed_rows(n);
    const int chunks = (k_packed + CHUNK_K - 1) / CHUNK_K;
    return (long long)chunks * 2 * n_pad * 128;
}

KDPC_API int kdpc_pack_weight(int n, int k_src, int mode, int d, int wn, const float *w, void *out,
                              kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(w && out && n > 0 && k_src > 0);
    if (n > 32768 || (n > 256 && mode != 0)) return KDPC_EUNSUPPORTED;
    int k_packed = k_src;
    if (mode == 1) {
        KDPC_CHECK_ARGS(d >= 0 && wn > 0 && k_src == (d + 3) * wn);
        k_packed = (d + 4) * wn;
    }
    if ((reinterpret_cast<uintptr_t>(out) % 16) != 0) return KDPC_EINVAL;
    const int n_pad = packed_rows(n);
    const int chunks = (k_packed + CHUNK_K - 1) / CHUNK_K;
    const long long total = (long long)chunks * n_pad * 8;
    pack_weight_kernel<<<(unsigned)div_up_ll(total, 256), 256, 0, to_stream(stream)>>>(
        n, k_src, n_pad, chunks, mode, d, wn, w, reinterpret_cast<unsigned char *>(out));
    KDPC_RETURN_LAST();
}

KDPC_API long long kdpc_linear_tc_ws_bytes(long long m, int n, int k) {
    if (m <= 0 || n <= 0 || n > 256 || k <= 0) return 0;          // (wide layers: column blocks, no workspace)
    GemmShape g = make_shape(m, n, k, nullptr);
    if (!(kdpc_linear_split_n && plan_split_n(g))) plan_split_k(g);
    return (long long)split_k_ws_bytes(g);
}

KDPC_API int kdpc_linear_tc(long long m, int n, int k, const float *x, int ldx, const void *wpacked,
                            const float *scale, const float *shift, float slope, float clamp_lo, float clamp_hi,
                            const float *residual, void *ws, float *out, int ldo, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(x && wpacked && out && m > 0 && n > 0 && k > 0 && ldx >= k && ldo >= n);
    if (n > 32768) return KDPC_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) % 16) != 0 || (reinterpret_cast<uintptr_t>(out) % 16) != 0 ||
        (reinterpret_cast<uintptr_t>(wpacked) % 16) != 0 || (reinterpret_cast<uintptr_t>(ws) % 16) != 0)
        return KDPC_EINVAL;
    const bool wide = n > 256;
    // WIDE layers (n > 256): ONE launch over (row tile, 128-column block) work items against a weight packed in whole column
    // blocks - instead of one pack + one launch per 256 columns from the host (the KD step issued 257 such launches)
    auto widen = [&](GemmShape &gs) {
        gs.n = n;
        gs.w_n_pad = packed_rows(n);
        gs.splits = gs.w_n_pad / 128;
        gs.nsplit = 1;
        gs.chunks_per_split = gs.num_chunks;
        gs.n_pad = 128;
        gs.acc_stride = 128;
        gs.nacc_log2 = 2;
        gs.tmem_cols = 512;
        gs.stages = pick_stages(128, gs.raw_bytes * gs.raw_stages);
    };
    GemmShape g = make_shape(m, wide ? 128 : n, k, wpacked);
    if (wide) {
        if ((long long)g.num_tiles * (packed_rows(n) / 128) >= (1ll << 31)) return KDPC_EUNSUPPORTED;
        widen(g);
    } else if (!(kdpc_linear_split_n && plan_split_n(g)) && ws != nullptr) {
        plan_split_k(g);       // small-M layers: column blocks over idle SMs when there are >= 64 outputs, else (large K) split-K
    }
    if ((g.splits == 1 || g.nsplit) && (k & 7) == 0 && (ldx & 3) == 0 && m * (long long)TILE_M < (1ll << 31) && kdpc_tc_async_enabled() == 2) {
        // streaming layers, rows by 2-D tensor-map TMA: two UTMALDG per chunk from one thread.  Split-N plans too: every
        // K chunk of a (row tile, column block) item is in flight from the start instead of one register round trip per chunk
        using P = PlainTmaProducer;
        P::Args pa;
        if (make_row_tensor_map(&pa.tmap, x, m, k, ldx)) {
            pa.k = k;
            for (int raw = P::kLookahead + 1; raw >= 2; --raw) {
                GemmShape ga = make_shape(m, wide ? 128 : n, k, wpacked, P::kRawBytes, raw);
                if (wide) widen(ga);
                else if (g.nsplit && !plan_split_n(ga)) continue;
                if (ga.stages < 2) continue;
                const size_t smem_a = smem_bytes(ga.n_pad, ga.stages, ga.raw_bytes * ga.raw_stages);
                auto kern_a = tc_gemm_kernel<P, StoreEpilogue>;
                KDPC_ENSURE_SMEM(kern_a, SMEM_BUDGET + 1024);
                StoreEpilogue::Args ea{scale, shift, slope, clamp_lo, clamp_hi, residual, out, ldo, nullptr};
                const long long work_a = ga.num_tiles * (ga.nsplit ? ga.splits : 1);
                const unsigned grid = (unsigned)(work_a < num_sms() ? work_a : num_sms());
                launch_tc(kern_a, grid, num_threads<P>(), smem_a, to_stream(stream), ga, pa, ea);
                KDPC_RETURN_LAST();
            }
        }
    }
    if (g.splits == 1 && (k & 7) == 0 && (ldx & 3) == 0 && m * (long long)TILE_M < (1ll << 31) && kdpc_tc_async_enabled()) {
        // streaming layers: asynchronous row fetch, 2 raw buffers ahead if they fit next to 2 operand stages, else 1
        using P = PlainAsyncProducer;
        for (int raw = P::kLookahead + 1; raw >= 2; --raw) {
            GemmShape ga = make_shape(m, n, k, wpacked, P::kRawBytes, raw);
            if (ga.stages < 2) continue;
            const size_t smem_a = smem_bytes(ga.n_pad, ga.stages, ga.raw_bytes * ga.raw_stages);
            auto kern_a = tc_gemm_kernel<P, StoreEpilogue>;
            KDPC_ENSURE_SMEM(kern_a, SMEM_BUDGET + 1024);
            P::Args pa{x, ldx, k};
            StoreEpilogue::Args ea{scale, shift, slope, clamp_lo, clamp_hi, residual, out, ldo, nullptr};
            const unsigned grid = (unsigned)(ga.num_tiles < num_sms() ? ga.num_tiles : num_sms());
            launch_tc(kern_a, grid, num_threads<P>(), smem_a, to_stream(stream), ga, pa, ea);
            KDPC_RETURN_LAST();
        }
    }
    const size_t smem = smem_bytes(g.n_pad, g.stages);
    auto kern = tc_gemm_kernel<PlainProducer, StoreEpilogue>;
    KDPC_ENSURE_SMEM(kern, 201 * 1024);
    PlainProducer::Args pa{x, ldx, k};
    StoreEpilogue::Args ea{scale, shift, slope, clamp_lo, clamp_hi, residual, out, ldo, reinterpret_cast<float *>(ws)};
    const long long work = g.num_tiles * g.splits;
    const unsigned grid = (unsigned)(work < num_sms() ? work : num_sms());
    launch_tc(kern, grid, num_threads<PlainProducer>(), smem, to_stream(stream), g, pa, ea);
    if (g.splits > 1 && !g.nsplit) return launch_splitk_reduce(g, ea, to_stream(stream));
    KDPC_RETURN_LAST();
}

KDPC_API int kdpc_linear_simt(long long m, int n, int k, const float *x, int ldx, const float *w,
                              const float *scale, const float *shift, float slope, float clamp_lo, float clamp_hi,
                              const float *residual, float *out, int ldo, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(x && w && out && m > 0 && n > 0 && k > 0 && ldx >= k && ldo >= n);
    cudaStream_t st = to_stream(stream);
    if (m >= 16384 && k <= 4 && (n & 3) == 0 && n <= 256 && residual == nullptr && (ldo & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const int lanes = 256 / (n >> 2);
        const long long want = (m + lanes - 1) / lanes;
        const unsigned grid = (unsigned)(want < 8LL * device_sms() ? want : 8LL * device_sms());
        linear_smallk_kernel<4><<<grid, 256, 0, st>>>(m, n, k, x, ldx, w, scale, shift, slope, clamp_lo, clamp_hi, out, ldo);
        KDPC_RETURN_LAST();
    }
    if (m >= 16384 && n <= 4 && (k & 3) == 0 && (ldx & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0) {
        const long long want = (m + 31) / 32;
        const unsigned grid = (unsigned)(want < 8LL * device_sms() ? want : 8LL * device_sms());
        linear_smalln_kernel<4><<<grid, 256, 0, st>>>(m, n, k, x, ldx, w, scale, shift, slope, clamp_lo, clamp_hi, residual, out, ldo);
        KDPC_RETURN_LAST();
    }
    const long long total = m * ((n + 3) / 4);
    linear_simt_kernel<<<(unsigned)div_up_ll(total, 256), 256, 0, to_stream(stream)>>>(
        m, n, k, x, ldx, w, scale, shift, slope, clamp_lo, clamp_hi, residual, out, ldo);
    KDPC_RETURN_LAST();
}
