// Backward of the fused cost-volume half (CrossLayerLight.cross, pointconv_util.py:1826-1850 through loss.backward(),
// distilTrain.py:180) for the 8192-point level (D = D' = 32, K = 32), exploiting what the max over the K neighbours does
// to the gradient.
//
//   forward (costvol_tc.cu):  out[i, c] = act2( max_k  z[i, k, c] ),   z[i, k, :] = W h[i, k, :] + b,
//                             h[i, k, :] = act1( p2q[idx[i, k], :] + p1q[i, :] )
//
// autograd over the unfused op chain runs DENSE over all B*N*K = 2 M rows: three 268 MB tensors forward, the same again
// backward (dX, dW, two masks, a scatter) although the max lets the gradient of out[i, c] through to ONE neighbour
// k*(i, c) only: 32 of the 1024 (k, c) entries of a point.  Here one warp owns a point (lane = neighbour):
//   1. recompute h (registers, 32 channels per lane) and, channel by channel, z and its arg-max over the lanes
//      (redux.sync.max on order-preserving ints + ballot: the first maximal neighbour, like torch.max);
//   2. g = dout[i, c] * act2'(max): the ONE lane k* adds g W[c, :] to its dh; lane d adds g h[k*, d] to its column of the
//      weight gradient (h staged in shared memory), lane c adds g to the bias gradient;
//   3. da = dh * act1'(h) goes out as a dense [B*N*K, 32] row block for the deterministic CSR scatter into dp2q
//      (scatter.cu; no float atomics anywhere), and summed over the lanes in lane order into dp1q[i, :].
// Weight / bias gradients: per-warp register accumulators over the warp's points (fixed assignment), per-warp partials
// to a workspace, summed in warp order by costvol_grad_reduce_kernel: run-to-run identical.
#include "common.cuh"

namespace kdpc {

constexpr int CG_D = 32, CG_K = 32, CG_WARPS = 4, CG_HP = 33;      // h rows padded to 33 floats: conflict-free both ways
constexpr int CG_PART = CG_D * CG_D + CG_D;                        // floats of one warp's partial (dW, db)

__device__ __forceinline__ int cg_f2ord(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float cg_ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

__global__ void __launch_bounds__(CG_WARPS * 32)
costvol_grad32_kernel(long long points, int s, int n, const float *__restrict__ p1q, const float *__restrict__ p2q,
                      const int *__restrict__ idx, const float *__restrict__ w, const float *__restrict__ bias, float slope_pre,
                      float slope_post, const float *__restrict__ gout, float *__restrict__ g_p1q, float *__restrict__ g_rows,
                      float *__restrict__ partial) {
    __shared__ __align__(16) float sw[CG_D * CG_D];               // W [c][d]
    __shared__ float sb[CG_D];
    __shared__ float sh[CG_WARPS][CG_K * CG_HP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < CG_D * CG_D; i += blockDim.x) sw[i] = __ldg(w + i);
    if (threadIdx.x < CG_D) sb[threadIdx.x] = bias != nullptr ? __ldg(bias + threadIdx.x) : 0.f;
    __syncthreads();
    float *hs = sh[warp];
    float dw_acc[CG_D];                                           // lane d: dW[c][d], c = 0..31
#pragma unroll
    for (int c = 0; c < CG_D; ++c) dw_acc[c] = 0.f;
    float db_acc = 0.f;                                           // lane c: db[c]
    const long long nwarps = (long long)gridDim.x * CG_WARPS, gw = (long long)blockIdx.x * CG_WARPS + warp;
    for (long long pt = gw; pt < points; pt += nwarps) {
        const long long b = pt / s;
        const int j = __ldg(idx + pt * CG_K + lane);
        const float4 *r2 = reinterpret_cast<const float4 *>(p2q + (b * n + j) * CG_D);
        const float4 *r1 = reinterpret_cast<const float4 *>(p1q + pt * CG_D);
        float h[CG_D];
#pragma unroll
        for (int q = 0; q < CG_D / 4; ++q) {
            const float4 a2 = __ldg(r2 + q), a1 = __ldg(r1 + q);
            const float v[4] = {a2.x + a1.x, a2.y + a1.y, a2.z + a1.z, a2.w + a1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) h[4 * q + e] = fmaxf(v[e], v[e] * slope_pre);       // leaky / ReLU (0 <= slope < 1)
        }
        const float gl = __ldg(gout + pt * CG_D + lane);          // lane c holds dout[i, c]
#pragma unroll
        for (int d = 0; d < CG_D; ++d) hs[lane * CG_HP + d] = h[d];
        __syncwarp();
        float dh[CG_D];
#pragma unroll
        for (int d = 0; d < CG_D; ++d) dh[d] = 0.f;
#pragma unroll
        for (int c = 0; c < CG_D; ++c) {
            float wr[CG_D];
#pragma unroll
            for (int q = 0; q < CG_D / 4; ++q) {
                const float4 t = *reinterpret_cast<const float4 *>(sw + c * CG_D + 4 * q);   // broadcast
                wr[4 * q] = t.x; wr[4 * q + 1] = t.y; wr[4 * q + 2] = t.z; wr[4 * q + 3] = t.w;
            }
            float z = sb[c];
#pragma unroll
            for (int d = 0; d < CG_D; ++d) z = fmaf(wr[d], h[d], z);
            const int zo = cg_f2ord(z);
            const int m = __reduce_max_sync(0xffffffffu, zo);
            const int ks = __ffs(__ballot_sync(0xffffffffu, zo == m)) - 1;                   // first maximal neighbour
            float g = __shfl_sync(0xffffffffu, gl, c);
            g = cg_ord2f(m) > 0.f ? g : g * slope_post;
            if (lane == ks) {
#pragma unroll
                for (int d = 0; d < CG_D; ++d) dh[d] = fmaf(g, wr[d], dh[d]);
            }
            dw_acc[c] = fmaf(g, hs[ks * CG_HP + lane], dw_acc[c]);
            if (lane == c) db_acc += g;
        }
        __syncwarp();
        // da = dh * act1'(h): the row of the scatter into dp2q, and (summed over the neighbours in lane order) dp1q
        float4 *orow = reinterpret_cast<float4 *>(g_rows + (pt * CG_K + lane) * CG_D);
#pragma unroll
        for (int q = 0; q < CG_D / 4; ++q) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d = 4 * q + e;
                o[e] = h[d] > 0.f ? dh[d] : dh[d] * slope_pre;
                hs[lane * CG_HP + d] = o[e];
            }
            orow[q] = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
        float sum = 0.f;
#pragma unroll
        for (int kk = 0; kk < CG_K; ++kk) sum += hs[kk * CG_HP + lane];
        g_p1q[pt * CG_D + lane] = sum;
        __syncwarp();
    }
    float *pw = partial + gw * CG_PART;
#pragma unroll
    for (int c = 0; c < CG_D; ++c) pw[c * CG_D + lane] = dw_acc[c];
    pw[CG_D * CG_D + lane] = db_acc;
}

// dW[c][d] / db[c] = sum over the warps' partials in warp order (thread = one output, four running sums over interleaved
// warps added in a fixed order)
__global__ void __launch_bounds__(256)
costvol_grad_reduce_kernel(int nwarps, const float *__restrict__ partial, float *__restrict__ g_w, float *__restrict__ g_b) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= CG_PART) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int wi = 0;
    for (; wi + 3 < nwarps; wi += 4) {
        a0 += partial[(size_t)wi * CG_PART + o];
        a1 += partial[(size_t)(wi + 1) * CG_PART + o];
        a2 += partial[(size_t)(wi + 2) * CG_PART + o];
        a3 += partial[(size_t)(wi + 3) * CG_PART + o];
    }
    for (; wi < nwarps; ++wi) a0 += partial[(size_t)wi * CG_PART + o];
    const float t = (a0 + a1) + (a2 + a3);
    if (o < CG_D * CG_D) g_w[o] = t;
    else if (g_b != nullptr) g_b[o - CG_D * CG_D] = t;
}

static int costvol_grad_grid() { return 3 * device_sms(); }       // 165 registers x 128 threads: three CTAs per SM

}  // namespace kdpc

using namespace kdpc;

KDPC_API long long kdpc_costvol_grad_ws_bytes(void) {
    return (long long)costvol_grad_grid() * CG_WARPS * CG_PART * (long long)sizeof(float);
}

/* Gradients of out = kdpc_costvol_fused(...) in its folded form (p1q = points1 + pos_b - pos_w xyz1, p2q = points2 + pos_w xyz2):
 * grad_p1q [b,s,d], grad_rows [b*s*k, d] (row (i, k) = gradient of the row gathered from p2q[idx[i, k]]: scatter it with
 * kdpc_scatter_rows_csr), grad_w [d_out, d], grad_b [d_out] or NULL.  d = d_out = k = 32.  w: fp32 [d_out, d]. */
KDPC_API int kdpc_costvol_grad(int b, int s, int n, int k, int d, int d_out, const float *p1q, const float *p2q, const int *idx,
                               const float *w, const float *bias, float slope_pre, float slope_post, const float *grad_out,
                               void *ws, float *grad_p1q, float *grad_rows, float *grad_w, float *grad_b, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(p1q && p2q && idx && w && grad_out && ws && grad_p1q && grad_rows && grad_w && b > 0 && s > 0 && n > 0);
    if (k != CG_K || d != CG_D || d_out != CG_D || slope_pre < 0.f || slope_pre >= 1.f) return KDPC_EUNSUPPORTED;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p1q) | reinterpret_cast<uintptr_t>(p2q) | reinterpret_cast<uintptr_t>(grad_rows) |
                         reinterpret_cast<uintptr_t>(ws);
    if (al % 16 != 0) return KDPC_EINVAL;
    cudaStream_t st = to_stream(stream);
    const int grid = costvol_grad_grid();
    costvol_grad32_kernel<<<grid, CG_WARPS * 32, 0, st>>>((long long)b * s, s, n, p1q, p2q, idx, w, bias, slope_pre, slope_post,
                                                         grad_out, grad_p1q, grad_rows, reinterpret_cast<float *>(ws));
    costvol_grad_reduce_kernel<<<(CG_PART + 255) / 256, 256, 0, st>>>(grid * CG_WARPS, reinterpret_cast<const float *>(ws), grad_w, grad_b);
    KDPC_RETURN_LAST();
}
