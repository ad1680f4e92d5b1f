// PointConv pieces for sm_100a (unfused building blocks).
//
//  * weightnet : the per-neighbour MLP 3 -> h1 -> h2 -> wout with ReLU after EVERY layer
//                (reference WeightNet, pointconv_util.py:184-215, bn=False), which the reference
//                runs as three cuDNN 1x1 convolutions on a strided [B,3,K,S] view.
//  * pointconv_agg : out[r, c*wout + w] = sum_k grouped[r,k,c] * wn[r,k,w]
//                (the torch.matmul at pointconv_util.py:249 / :437, B*S tiny matrices).
// The fully fused PointConv (gather + weightnet + aggregation + Linear on tcgen05) lives in
// pointconv_fused.cu; these kernels serve the stand-alone WeightNet / group API and the backward.
#include "common.cuh"

namespace kdpc {

template <int H1, int H2, int WOUT>
__global__ void __launch_bounds__(256)
weightnet_kernel(long long rows, const float *__restrict__ in, int in_stride, const float *__restrict__ w1,
                 const float *__restrict__ b1, const float *__restrict__ w2, const float *__restrict__ b2,
                 const float *__restrict__ w3, const float *__restrict__ b3, float *__restrict__ out) {
    __shared__ float sw1[H1 * 3], sb1[H1], sw2[H2 * H1], sb2[H2], sw3[WOUT * H2], sb3[WOUT];
    for (int i = threadIdx.x; i < H1 * 3; i += blockDim.x) sw1[i] = w1[i];
    for (int i = threadIdx.x; i < H1; i += blockDim.x) sb1[i] = b1[i];
    for (int i = threadIdx.x; i < H2 * H1; i += blockDim.x) sw2[i] = w2[i];
    for (int i = threadIdx.x; i < H2; i += blockDim.x) sb2[i] = b2[i];
    for (int i = threadIdx.x; i < WOUT * H2; i += blockDim.x) sw3[i] = w3[i];
    for (int i = threadIdx.x; i < WOUT; i += blockDim.x) sb3[i] = b3[i];
    __syncthreads();
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *x = in + r * in_stride;
    const float x0 = x[0], x1 = x[1], x2 = x[2];
    float h1[H1], h2[H2];
#pragma unroll
    for (int o = 0; o < H1; ++o)
        h1[o] = fmaxf(sb1[o] + sw1[o * 3 + 0] * x0 + sw1[o * 3 + 1] * x1 + sw1[o * 3 + 2] * x2, 0.f);
#pragma unroll
    for (int o = 0; o < H2; ++o) {
        float a = sb2[o];
#pragma unroll
        for (int i = 0; i < H1; ++i) a += sw2[o * H1 + i] * h1[i];
        h2[o] = fmaxf(a, 0.f);
    }
    float *op = out + r * WOUT;
#pragma unroll
    for (int o4 = 0; o4 < WOUT; o4 += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float a = sb3[o4 + u];
#pragma unroll
            for (int i = 0; i < H2; ++i) a += sw3[(o4 + u) * H2 + i] * h2[i];
            v[u] = fmaxf(a, 0.f);
        }
        *reinterpret_cast<float4 *>(op + o4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// One CTA per AGG_ROWS points; thread = one channel c, WOUT accumulators.
// grouped[r,k,c] is read coalesced across c; wn[r,k,:] comes from shared memory (broadcast).
constexpr int AGG_THREADS = 128;

template <int WOUT>
__global__ void __launch_bounds__(AGG_THREADS)
pointconv_agg_kernel(long long rows, int k, int c, const float *__restrict__ grouped, const float *__restrict__ wn,
                     float *__restrict__ out) {
    extern __shared__ float swn[];                        // [k][WOUT]
    const long long r = blockIdx.x;
    const float *wr = wn + r * (long long)k * WOUT;
    for (int i = threadIdx.x; i < k * WOUT; i += AGG_THREADS) swn[i] = wr[i];
    __syncthreads();
    const float *gr = grouped + r * (long long)k * c;
    float *orow = out + r * (long long)c * WOUT;
    for (int ci = threadIdx.x; ci < c; ci += AGG_THREADS) {
        float acc[WOUT];
#pragma unroll
        for (int w = 0; w < WOUT; ++w) acc[w] = 0.f;
        for (int kk = 0; kk < k; ++kk) {
            const float g = __ldg(gr + (size_t)kk * c + ci);
            const float *wk = swn + kk * WOUT;
#pragma unroll
            for (int w = 0; w < WOUT; ++w) acc[w] = fmaf(g, wk[w], acc[w]);
        }
        float4 *o4 = reinterpret_cast<float4 *>(orow + (size_t)ci * WOUT);
#pragma unroll
        for (int w = 0; w < WOUT; w += 4) st_stream_f4(o4 + (w >> 2), make_float4(acc[w], acc[w + 1], acc[w + 2], acc[w + 3]));
    }
}

// Backward of the aggregation (autograd of pointconv_util.py:249, two bmm of B*S tiny matrices in torch):
//   g_grouped[r,k,c] = sum_w wn[r,k,w] * g[r,c,w]          g_wn[r,k,w] = sum_c grouped[r,k,c] * g[r,c,w]
// One CTA per point r, thread = channel c (same mapping as the forward): it reads its 64-byte slice g[r,c,:] once,
// produces g_grouped[r,:,c] (coalesced across c) and adds grouped[r,k,c] * g[r,c,:] into per-thread partial sums of
// g_wn, which a fixed shuffle + shared-memory tree then reduces over the channels (deterministic).
template <int WOUT, int KCH>
__global__ void __launch_bounds__(AGG_THREADS)
pointconv_agg_grad_kernel(long long rows, int k, int c, const float *__restrict__ grouped, const float *__restrict__ wn,
                          const float *__restrict__ g, float *__restrict__ g_grouped, float *__restrict__ g_wn) {
    extern __shared__ float swn[];                            // [k][WOUT]
    __shared__ __align__(16) float sred[AGG_THREADS / 32][KCH * WOUT];
    const long long r = blockIdx.x;
    const float *wr = wn + r * (long long)k * WOUT;
    for (int i = threadIdx.x; i < k * WOUT; i += AGG_THREADS) swn[i] = wr[i];
    __syncthreads();
    const float *gr = grouped + r * (long long)k * c;
    const float *grow = g + r * (long long)c * WOUT;
    float *ogr = g_grouped != nullptr ? g_grouped + r * (long long)k * c : nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k0 = 0; k0 < k; k0 += KCH) {                     // neighbours in chunks of KCH: KCH x WOUT partial sums per thread
        float part[KCH][WOUT];
#pragma unroll
        for (int kk = 0; kk < KCH; ++kk)
#pragma unroll
            for (int w = 0; w < WOUT; ++w) part[kk][w] = 0.f;
        for (int ci = threadIdx.x; ci < c; ci += AGG_THREADS) {
            float gv[WOUT];
            const float4 *g4 = reinterpret_cast<const float4 *>(grow + (size_t)ci * WOUT);
#pragma unroll
            for (int w = 0; w < WOUT; w += 4) {
                const float4 t = __ldg(g4 + (w >> 2));
                gv[w] = t.x; gv[w + 1] = t.y; gv[w + 2] = t.z; gv[w + 3] = t.w;
            }
#pragma unroll
            for (int kk = 0; kk < KCH; ++kk) {
                if (k0 + kk < k) {
                    if (ogr != nullptr) {
                        float a = 0.f;
#pragma unroll
                        for (int w = 0; w < WOUT; ++w) a = fmaf(swn[(k0 + kk) * WOUT + w], gv[w], a);
                        ogr[(size_t)(k0 + kk) * c + ci] = a;
                    }
                    if (g_wn != nullptr) {
                        const float x = __ldg(gr + (size_t)(k0 + kk) * c + ci);
#pragma unroll
                        for (int w = 0; w < WOUT; ++w) part[kk][w] = fmaf(x, gv[w], part[kk][w]);
                    }
                }
            }
        }
        if (g_wn == nullptr) continue;
        // warp-level sum of the KCH x WOUT partials.  The first 128 of them by recursive halving (reduce-scatter): at lane
        // distance 16, 8, .. 1 every lane keeps the half that matches its lane bit and adds its partner's copy of it -
        // 124 shuffles, lane l ends up with the warp sums of values 4l .. 4l+3 (a butterfly all-reduce of every value was
        // 5 shuffles each: 720 per row, more than the arithmetic).  Values beyond 128 (KCH = 9): plain butterfly.
        static_assert(WOUT == 16 && (KCH == 8 || KCH == 9), "reduction layout");
        {
            float *v = &part[0][0];
#pragma unroll
            for (int m = 16, n = 64; m >= 1; m >>= 1, n >>= 1) {
                const bool up = (lane & m) != 0;
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    const float keep = up ? v[j + n] : v[j];
                    const float give = up ? v[j] : v[j + n];
                    v[j] = keep + __shfl_xor_sync(0xffffffffu, give, m);
                }
            }
            *reinterpret_cast<float4 *>(&sred[warp][4 * lane]) = make_float4(v[0], v[1], v[2], v[3]);
            if (KCH == 9) {
#pragma unroll
                for (int w = 0; w < WOUT; ++w) {
                    float t = part[KCH - 1][w];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    if (lane == 0) sred[warp][128 + w] = t;
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < KCH * WOUT; i += AGG_THREADS) {
            if (k0 + i / WOUT < k) {
                float v = 0.f;
#pragma unroll
                for (int wp = 0; wp < AGG_THREADS / 32; ++wp) v += sred[wp][i];
                g_wn[r * (long long)k * WOUT + (size_t)k0 * WOUT + i] = v;
            }
        }
        __syncthreads();
    }
}


// The same backward with ONE WARP per point and channel blocks of 64 (k <= 16): the CTA-per-point kernel above keeps
// k x 16 partial sums per thread (248 registers: two CTAs = 8 warps per SM, every CTA a short latency chain, half its
// threads idle in the second pass over c = 131 channels) and ran 10x above both its HBM and its FMA bound (1.9 ms for
// 65536 x 9 x 131).  Here, per block of 64 channels:
//   phase 1  lane = channel (two per lane): its 64-byte slice g[r,c,:] goes to registers (fully coalesced across the
//            warp) and to shared memory, grouped[r,:,c] to shared memory; g_grouped[r,kk,c] = <wn[r,kk,:], g[r,c,:]> with
//            the wn rows broadcast from shared memory;
//   phase 2  lane = (w, kk parity): g_wn[r,kk,w] += sum_c grouped[r,kk,c] * g[r,c,w] from shared memory, one
//            accumulator per (kk, w) walking c in ascending order (deterministic, no cross-lane reduction at all).
// 80-112 registers, 4 warps and 33-40 KB of shared memory per CTA: 5-6 CTAs per SM.
constexpr int AGW_WARPS = 4, AGW_CB = 64, AGW_GP = 20;        // g rows padded to 20 floats: conflict-free 16-byte accesses by row

template <int KH>                                             // k <= 2 * KH
__global__ void __launch_bounds__(AGW_WARPS * 32)
pointconv_agg_grad_warp_kernel(long long rows, int k, int c, const float *__restrict__ grouped, const float *__restrict__ wn,
                               const float *__restrict__ g, float *__restrict__ g_grouped, float *__restrict__ g_wn) {
    __shared__ __align__(16) float s_g[AGW_WARPS][AGW_CB * AGW_GP];
    __shared__ __align__(16) float s_x[AGW_WARPS][2 * KH * AGW_CB];
    __shared__ __align__(16) float s_w[AGW_WARPS][2 * KH * 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * AGW_WARPS + warp;
    if (r >= rows) return;                                    // (warps never meet at a CTA barrier)
    float *sg = s_g[warp], *sx = s_x[warp], *sw = s_w[warp];
    {
        const float4 *w4 = reinterpret_cast<const float4 *>(wn + r * (long long)k * 16);
        for (int i = lane; i < k * 4; i += 32) reinterpret_cast<float4 *>(sw)[i] = __ldg(w4 + i);
    }
    __syncwarp();
    const int w = lane & 15, kh = lane >> 4;
    float acc[KH];
#pragma unroll
    for (int i = 0; i < KH; ++i) acc[i] = 0.f;
    const float *xrow = grouped + r * (long long)k * c;
    float *orow = g_grouped != nullptr ? g_grouped + r * (long long)k * c : nullptr;
    for (int c0 = 0; c0 < c; c0 += AGW_CB) {
        const int nv = min(AGW_CB, c - c0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int cl = lane + 32 * h;
            float4 *sgr = reinterpret_cast<float4 *>(sg + cl * AGW_GP);
            if (cl < nv) {
                const float4 *gp = reinterpret_cast<const float4 *>(g + (r * (long long)c + c0 + cl) * 16);
                const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2), g3 = __ldg(gp + 3);
                sgr[0] = g0; sgr[1] = g1; sgr[2] = g2; sgr[3] = g3;
                if (g_wn != nullptr) {
                    float xr[2 * KH];
#pragma unroll
                    for (int kk = 0; kk < 2 * KH; ++kk) xr[kk] = kk < k ? __ldg(xrow + (size_t)kk * c + c0 + cl) : 0.f;
#pragma unroll
                    for (int kk = 0; kk < 2 * KH; ++kk)
                        if (kk < k) sx[kk * AGW_CB + cl] = xr[kk];
                }
                if (orow != nullptr) {
                    const float2 ga[8] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w),
                                          make_float2(g2.x, g2.y), make_float2(g2.z, g2.w), make_float2(g3.x, g3.y), make_float2(g3.z, g3.w)};
                    for (int kk = 0; kk < k; ++kk) {
                        const float4 *wr = reinterpret_cast<const float4 *>(sw + kk * 16);
                        const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
                        float2 a = make_float2(0.f, 0.f);       // even / odd w, added at the end (packed fp32 FMAs)
                        a = __ffma2_rn(make_float2(w0.x, w0.y), ga[0], a); a = __ffma2_rn(make_float2(w0.z, w0.w), ga[1], a);
                        a = __ffma2_rn(make_float2(w1.x, w1.y), ga[2], a); a = __ffma2_rn(make_float2(w1.z, w1.w), ga[3], a);
                        a = __ffma2_rn(make_float2(w2.x, w2.y), ga[4], a); a = __ffma2_rn(make_float2(w2.z, w2.w), ga[5], a);
                        a = __ffma2_rn(make_float2(w3.x, w3.y), ga[6], a); a = __ffma2_rn(make_float2(w3.z, w3.w), ga[7], a);
                        orow[(size_t)kk * c + c0 + cl] = a.x + a.y;
                    }
                }
            } else if (g_wn != nullptr && cl < ((nv + 3) & ~3)) {     // zero padding up to the next multiple of 4 channels (phase 2 walks by 4)
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                sgr[0] = z; sgr[1] = z; sgr[2] = z; sgr[3] = z;
                for (int kk = 0; kk < k; ++kk) sx[kk * AGW_CB + cl] = 0.f;
            }
        }
        if (g_wn == nullptr) continue;
        __syncwarp();
        for (int cc = 0; cc < nv; cc += 4) {
            const float gv0 = sg[(cc + 0) * AGW_GP + w], gv1 = sg[(cc + 1) * AGW_GP + w];
            const float gv2 = sg[(cc + 2) * AGW_GP + w], gv3 = sg[(cc + 3) * AGW_GP + w];
#pragma unroll
            for (int i = 0; i < KH; ++i) {
                const int kk = 2 * i + kh;
                if (kk < k) {
                    const float4 xv = *reinterpret_cast<const float4 *>(sx + kk * AGW_CB + cc);
                    acc[i] = fmaf(xv.x, gv0, acc[i]);
                    acc[i] = fmaf(xv.y, gv1, acc[i]);
                    acc[i] = fmaf(xv.z, gv2, acc[i]);
                    acc[i] = fmaf(xv.w, gv3, acc[i]);
                }
            }
        }
        __syncwarp();
    }
    if (g_wn != nullptr) {
#pragma unroll
        for (int i = 0; i < KH; ++i) {
            const int kk = 2 * i + kh;
            if (kk < k) g_wn[(r * (long long)k + kk) * 16 + w] = acc[i];
        }
    }
}

}  // namespace kdpc

using namespace kdpc;

KDPC_API int kdpc_pointconv_agg_grad(long long rows, int k, int c, int wout, const float *grouped, const float *wn,
                                     const float *grad_out, float *grad_grouped, float *grad_wn, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(grouped && wn && grad_out && (grad_grouped || grad_wn) && rows > 0 && k > 0 && c > 0);
    if ((reinterpret_cast<uintptr_t>(grad_out) % 16) != 0 || rows > 0x7fffffffLL) return KDPC_EINVAL;
    if (wout != 16) return KDPC_EUNSUPPORTED;
    const size_t smem = (size_t)k * wout * sizeof(float);
    if (smem > 32 * 1024) return KDPC_EUNSUPPORTED;
    cudaStream_t st = to_stream(stream);
    if (k <= 16 && (reinterpret_cast<uintptr_t>(wn) % 16) == 0) {
        const unsigned grid = (unsigned)div_up_ll(rows, AGW_WARPS);
        if (k <= 10) pointconv_agg_grad_warp_kernel<5><<<grid, AGW_WARPS * 32, 0, st>>>(rows, k, c, grouped, wn, grad_out, grad_grouped, grad_wn);
        else pointconv_agg_grad_warp_kernel<8><<<grid, AGW_WARPS * 32, 0, st>>>(rows, k, c, grouped, wn, grad_out, grad_grouped, grad_wn);
        KDPC_RETURN_LAST();
    }
    if (k <= 9)
        pointconv_agg_grad_kernel<16, 9><<<(unsigned)rows, AGG_THREADS, smem, st>>>(rows, k, c, grouped, wn, grad_out, grad_grouped, grad_wn);
    else
        pointconv_agg_grad_kernel<16, 8><<<(unsigned)rows, AGG_THREADS, smem, st>>>(rows, k, c, grouped, wn, grad_out, grad_grouped, grad_wn);
    KDPC_RETURN_LAST();
}

KDPC_API int kdpc_weightnet(long long rows, const float *in, int in_stride, int h1, int h2, int wout,
                            const float *w1, const float *b1, const float *w2, const float *b2,
                            const float *w3, const float *b3, float *out, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(in && w1 && b1 && w2 && b2 && w3 && b3 && out && rows > 0 && in_stride >= 3);
    if ((reinterpret_cast<uintptr_t>(out) % 16) != 0) return KDPC_EINVAL;
    cudaStream_t st = to_stream(stream);
    const unsigned grid = (unsigned)div_up_ll(rows, 256);
#define KDPC_WN_CASE(A, B_, C) \
    if (h1 == A && h2 == B_ && wout == C) { \
        weightnet_kernel<A, B_, C><<<grid, 256, 0, st>>>(rows, in, in_stride, w1, b1, w2, b2, w3, b3, out); \
        return (int)cudaGetLastError(); }
    KDPC_WN_CASE(8, 8, 4)
    KDPC_WN_CASE(8, 8, 8)
    KDPC_WN_CASE(8, 8, 16)
    KDPC_WN_CASE(8, 8, 32)
    KDPC_WN_CASE(8, 8, 48)
#undef KDPC_WN_CASE
    return KDPC_EUNSUPPORTED;
}

KDPC_API int kdpc_pointconv_agg(long long rows, int k, int c, int wout, const float *grouped, const float *wn,
                                float *out, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(grouped && wn && out && rows > 0 && k > 0 && c > 0);
    if ((reinterpret_cast<uintptr_t>(out) % 16) != 0 || rows > 0x7fffffffLL) return KDPC_EINVAL;
    cudaStream_t st = to_stream(stream);
    const size_t smem = (size_t)k * wout * sizeof(float);
    if (smem > 48 * 1024) return KDPC_EUNSUPPORTED;
#define KDPC_AGG_CASE(W) \
    if (wout == W) { \
        pointconv_agg_kernel<W><<<(unsigned)rows, AGG_THREADS, smem, st>>>(rows, k, c, grouped, wn, out); \
        return (int)cudaGetLastError(); }
    KDPC_AGG_CASE(4)
    KDPC_AGG_CASE(8)
    KDPC_AGG_CASE(16)
    KDPC_AGG_CASE(32)
    KDPC_AGG_CASE(48)
#undef KDPC_AGG_CASE
    return KDPC_EUNSUPPORTED;
}
