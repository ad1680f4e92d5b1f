// Backward of the fused cost-volume half (CrossLayerLight.cross, pointconv_util.py:1826-1850 through loss.backward(),
// distilTrain.py:180) for the 8192- and 2048-point levels (K = 32, D = D' = 32 or 64), exploiting what the max over the K
// neighbours does to the gradient.
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

constexpr int CG_K = 32, CG_WARPS = 4;

__device__ __forceinline__ int cg_f2ord(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float cg_ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// D = D' = 32 (8192-point level): weight-gradient accumulators in registers (lane d: dW[0..31][d]).
// D = D' = 64 (2048-point level): 64 + 64 registers of h / dh per lane leave no room for them - every warp keeps its
// own [64][64] accumulator in shared memory instead (lane d adds to columns d and d + 32: conflict-free).
template <int D>
struct CgLayout {
    static constexpr int HP = D + 1;                               // h rows padded by one float: conflict-free both ways
    static constexpr int PART = D * D + D;                         // floats of one warp's partial (dW, db)
    static constexpr bool kSmemDw = D > 32;
    static constexpr int kWarpFloats = CG_K * HP + (kSmemDw ? D * D : 0);
    static constexpr size_t kSmemBytes = (size_t)(D * D + D + CG_WARPS * kWarpFloats) * sizeof(float);
};

template <int D>
__global__ void __launch_bounds__(CG_WARPS * 32)
costvol_grad_kernel(long long points, int s, int n, const float *__restrict__ p1q, const float *__restrict__ p2q,
                    const int *__restrict__ idx, const float *__restrict__ w, const float *__restrict__ bias, float slope_pre,
                    float slope_post, const float *__restrict__ gout, float *__restrict__ g_p1q, float *__restrict__ g_rows,
                    float *__restrict__ partial) {
    using L = CgLayout<D>;
    constexpr int HP = L::HP, NQ = D / 32;
    extern __shared__ __align__(16) float cg_smem[];
    float *sw = cg_smem;                                          // W [c][d]
    float *sb = sw + D * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *hs = sb + D + warp * L::kWarpFloats;                   // this warp's h rows [k][HP]
    float *dws = hs + CG_K * HP;                                  // (D > 32) this warp's dW accumulator [c][d]
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) sw[i] = __ldg(w + i);
    if (threadIdx.x < D) sb[threadIdx.x] = bias != nullptr ? __ldg(bias + threadIdx.x) : 0.f;
    if constexpr (L::kSmemDw)
        for (int i = lane; i < D * D; i += 32) dws[i] = 0.f;
    __syncthreads();
    float dw_acc[L::kSmemDw ? 1 : D];                             // (D = 32) lane d: dW[c][d], c = 0..31
#pragma unroll
    for (int c = 0; c < (L::kSmemDw ? 1 : D); ++c) dw_acc[c] = 0.f;
    float db_acc[NQ];                                             // lane l: db[l + 32 q]
#pragma unroll
    for (int q = 0; q < NQ; ++q) db_acc[q] = 0.f;
    const long long nwarps = (long long)gridDim.x * CG_WARPS, gw = (long long)blockIdx.x * CG_WARPS + warp;
    for (long long pt = gw; pt < points; pt += nwarps) {
        const long long b = pt / s;
        const int j = __ldg(idx + pt * CG_K + lane);
        const float4 *r2 = reinterpret_cast<const float4 *>(p2q + (b * n + j) * D);
        const float4 *r1 = reinterpret_cast<const float4 *>(p1q + pt * D);
        float h[D];
#pragma unroll
        for (int q = 0; q < D / 4; ++q) {
            const float4 a2 = __ldg(r2 + q), a1 = __ldg(r1 + q);
            const float v[4] = {a2.x + a1.x, a2.y + a1.y, a2.z + a1.z, a2.w + a1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) h[4 * q + e] = fmaxf(v[e], v[e] * slope_pre);       // leaky / ReLU (0 <= slope < 1)
        }
        float gl[NQ];                                             // lane l holds dout[i, l + 32 q]
#pragma unroll
        for (int q = 0; q < NQ; ++q) gl[q] = __ldg(gout + pt * D + 32 * q + lane);
#pragma unroll
        for (int d = 0; d < D; ++d) hs[lane * HP + d] = h[d];
        __syncwarp();
        float dh[D];
#pragma unroll
        for (int d = 0; d < D; ++d) dh[d] = 0.f;
        auto channel = [&](const int c) {
            // z[k, c] = b[c] + <W[c, :], h[k, :]> (weight row broadcast from shared memory)
            // four interleaved partial sums (a quarter of the chain latency) on packed fp32 FMAs (fma.rn.f32x2: half the issue slots)
            float2 za = make_float2(sb[c], 0.f), zb = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < D / 4; ++q) {
                const float4 t = *reinterpret_cast<const float4 *>(sw + c * D + 4 * q);
                za = __ffma2_rn(make_float2(t.x, t.y), make_float2(h[4 * q], h[4 * q + 1]), za);
                zb = __ffma2_rn(make_float2(t.z, t.w), make_float2(h[4 * q + 2], h[4 * q + 3]), zb);
            }
            const float z = (za.x + za.y) + (zb.x + zb.y);
            // (Measured and rejected: deferring the arg-max lane's dh += g W[c, :] until after the loop, each lane walking only
            // its own channels over a row-padded copy of W: -3 % at D = 32, and at D = 64 the copy costs the second resident CTA.)
            const int zo = cg_f2ord(z);
            const int m = __reduce_max_sync(0xffffffffu, zo);
            const int ks = __ffs(__ballot_sync(0xffffffffu, zo == m)) - 1;                   // first maximal neighbour
            float g = __shfl_sync(0xffffffffu, gl[c >> 5], c & 31);
            g = cg_ord2f(m) > 0.f ? g : g * slope_post;
            if (lane == ks) {
#pragma unroll
                for (int q = 0; q < D / 4; ++q) {
                    const float4 t = *reinterpret_cast<const float4 *>(sw + c * D + 4 * q);
                    const float2 g2 = make_float2(g, g);
                    const float2 d0 = __ffma2_rn(g2, make_float2(t.x, t.y), make_float2(dh[4 * q], dh[4 * q + 1]));
                    const float2 d1 = __ffma2_rn(g2, make_float2(t.z, t.w), make_float2(dh[4 * q + 2], dh[4 * q + 3]));
                    dh[4 * q] = d0.x; dh[4 * q + 1] = d0.y; dh[4 * q + 2] = d1.x; dh[4 * q + 3] = d1.y;
                }
            }
            if constexpr (L::kSmemDw) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) dws[c * D + 32 * q + lane] = fmaf(g, hs[ks * HP + 32 * q + lane], dws[c * D + 32 * q + lane]);
            } else {
                dw_acc[c] = fmaf(g, hs[ks * HP + lane], dw_acc[c]);
            }
            if (lane == (c & 31)) db_acc[c >> 5] += g;
        };
        if constexpr (L::kSmemDw) {
#pragma unroll 2
            for (int c0 = 0; c0 < D; c0 += 32) {                  // (c >> 5 stays a compile-time index of gl / db_acc)
#pragma unroll 4
                for (int c1 = 0; c1 < 32; ++c1) channel(c0 + c1);
            }
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c) channel(c);               // fully unrolled: dw_acc[c] lives in registers
        }
        __syncwarp();
        // da = dh * act1'(h): the row of the scatter into dp2q, and (summed over the neighbours in lane order) dp1q
        float4 *orow = reinterpret_cast<float4 *>(g_rows + (pt * CG_K + lane) * D);
#pragma unroll
        for (int q = 0; q < D / 4; ++q) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d = 4 * q + e;
                o[e] = h[d] > 0.f ? dh[d] : dh[d] * slope_pre;
                hs[lane * HP + d] = o[e];
            }
            orow[q] = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float sum = 0.f;
#pragma unroll
            for (int kk = 0; kk < CG_K; ++kk) sum += hs[kk * HP + 32 * q + lane];
            g_p1q[pt * D + 32 * q + lane] = sum;
        }
        __syncwarp();
    }
    float *pw = partial + gw * L::PART;
    if constexpr (L::kSmemDw) {
        for (int i = lane; i < D * D; i += 32) pw[i] = dws[i];
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) pw[c * D + lane] = dw_acc[c];
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) pw[D * D + 32 * q + lane] = db_acc[q];
}

// dW[c][d] / db[c] = sum over the warps' partials in warp order (thread = one output, four running sums over interleaved
// warps added in a fixed order)
__global__ void __launch_bounds__(256)
costvol_grad_reduce_kernel(int d, int nwarps, const float *__restrict__ partial, float *__restrict__ g_w, float *__restrict__ g_b) {
    const int part = d * d + d;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= part) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int wi = 0;
    for (; wi + 3 < nwarps; wi += 4) {
        a0 += partial[(size_t)wi * part + o];
        a1 += partial[(size_t)(wi + 1) * part + o];
        a2 += partial[(size_t)(wi + 2) * part + o];
        a3 += partial[(size_t)(wi + 3) * part + o];
    }
    for (; wi < nwarps; ++wi) a0 += partial[(size_t)wi * part + o];
    const float t = (a0 + a1) + (a2 + a3);
    if (o < d * d) g_w[o] = t;
    else if (g_b != nullptr) g_b[o - d * d] = t;
}

static int costvol_grad_grid(int d) { return (d <= 32 ? 3 : 2) * device_sms(); }       // CTAs per SM that fit (registers / shared memory)

}  // namespace kdpc

using namespace kdpc;

KDPC_API long long kdpc_costvol_grad_ws_bytes(void) {
    const long long a = (long long)costvol_grad_grid(32) * CG_WARPS * CgLayout<32>::PART;
    const long long b = (long long)costvol_grad_grid(64) * CG_WARPS * CgLayout<64>::PART;
    return (a > b ? a : b) * (long long)sizeof(float);
}

template <int D>
static int launch_costvol_grad(long long points, int s, int n, const float *p1q, const float *p2q, const int *idx, const float *w,
                               const float *bias, float slope_pre, float slope_post, const float *grad_out, float *ws,
                               float *grad_p1q, float *grad_rows, float *grad_w, float *grad_b, cudaStream_t st) {
    auto kern = costvol_grad_kernel<D>;
    KDPC_ENSURE_SMEM(kern, (int)CgLayout<D>::kSmemBytes);
    const int grid = costvol_grad_grid(D);
    kern<<<grid, CG_WARPS * 32, CgLayout<D>::kSmemBytes, st>>>(points, s, n, p1q, p2q, idx, w, bias, slope_pre, slope_post, grad_out,
                                                              grad_p1q, grad_rows, ws);
    costvol_grad_reduce_kernel<<<(CgLayout<D>::PART + 255) / 256, 256, 0, st>>>(D, grid * CG_WARPS, ws, grad_w, grad_b);
    return (int)cudaGetLastError();
}

/* Gradients of out = kdpc_costvol_fused(...) in its folded form (p1q = points1 + pos_b - pos_w xyz1, p2q = points2 + pos_w xyz2):
 * grad_p1q [b,s,d], grad_rows [b*s*k, d] (row (i, k) = gradient of the row gathered from p2q[idx[i, k]]: scatter it with
 * kdpc_scatter_rows_csr), grad_w [d_out, d], grad_b [d_out] or NULL.  k = 32, d = d_out = 32 or 64.  w: fp32 [d_out, d]. */
KDPC_API int kdpc_costvol_grad(int b, int s, int n, int k, int d, int d_out, const float *p1q, const float *p2q, const int *idx,
                               const float *w, const float *bias, float slope_pre, float slope_post, const float *grad_out,
                               void *ws, float *grad_p1q, float *grad_rows, float *grad_w, float *grad_b, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(p1q && p2q && idx && w && grad_out && ws && grad_p1q && grad_rows && grad_w && b > 0 && s > 0 && n > 0);
    if (k != CG_K || (d != 32 && d != 64) || d_out != d || slope_pre < 0.f || slope_pre >= 1.f) return KDPC_EUNSUPPORTED;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p1q) | reinterpret_cast<uintptr_t>(p2q) | reinterpret_cast<uintptr_t>(grad_rows) |
                         reinterpret_cast<uintptr_t>(ws);
    if (al % 16 != 0) return KDPC_EINVAL;
    cudaStream_t st = to_stream(stream);
    float *wsf = reinterpret_cast<float *>(ws);
    if (d == 32)
        return launch_costvol_grad<32>((long long)b * s, s, n, p1q, p2q, idx, w, bias, slope_pre, slope_post, grad_out, wsf, grad_p1q,
                                       grad_rows, grad_w, grad_b, st);
    return launch_costvol_grad<64>((long long)b * s, s, n, p1q, p2q, idx, w, bias, slope_pre, slope_post, grad_out, wsf, grad_p1q,
                                   grad_rows, grad_w, grad_b, st);
}
