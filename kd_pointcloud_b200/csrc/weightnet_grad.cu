// Backward of the WeightNet MLP 3 -> 8 -> 8 -> W (ReLU after every layer; reference pointconv_util.py:184-215, bn=False)
// in ONE kernel (training path of every PointConv / PointConvD: distilTrain.py:180 loss.backward()).
//
// The reference (and the unfused op chain) runs three 1x1 convolutions on a [B,3,K,S] view and lets autograd produce six
// skinny GEMMs + three ReLU masks over B*S*K rows (590 k rows at flow0): tiny channel counts, huge row counts - the
// cuBLAS SIMT kernels they land on were ~10 ms of the KD step.  Here a thread recomputes the forward of ITS row from the
// three coordinates (nothing but the input was saved), back-propagates the incoming gradient through the three layers in
// registers and parks (x, h1, h2, g1, g2, g3) in shared memory; the 32 lanes of the warp then each own <= 9 of the
// parameter gradients (one output row of dW3 / dW2 / dW1 plus its bias) and add the 32 parked rows in row order.
// Per-warp partials go to a workspace and weightnet_grad_reduce_kernel sums them in warp order: deterministic.
// Optionally writes the gradient w.r.t. the coordinates (g_in, stride 3).
#include "common.cuh"

namespace kdpc {

constexpr int WG_THREADS = 128;            // 4 warps: the parked rows (4 x 32 x 68 floats) fit the static 48 KB
constexpr int WG_REC = 68;                 // floats per parked row: x[0..7] | h1[8..15] | h2[16..23] | g1[24..31] | g2[32..39] | g3[40..40+W)

template <int WOUT>
__global__ void __launch_bounds__(WG_THREADS)
weightnet_grad_kernel(long long rows, const float *__restrict__ in, int in_stride, const float *__restrict__ g_out,
                      const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                      const float *__restrict__ b2, const float *__restrict__ w3, const float *__restrict__ b3,
                      float *__restrict__ partial, float *__restrict__ g_in) {
    static_assert(WOUT % 4 == 0 && WOUT <= 24, "lane ownership below covers up to 24 outputs");
    __shared__ float sw1[24], sb1[8], sw2[64], sb2[8], sw3[WOUT * 8], sb3[WOUT];
    __shared__ __align__(16) float rec[WG_THREADS / 32][32][WG_REC];
    for (int i = threadIdx.x; i < 24; i += blockDim.x) sw1[i] = w1[i];
    for (int i = threadIdx.x; i < 8; i += blockDim.x) { sb1[i] = b1[i]; sb2[i] = b2[i]; }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sw2[i] = w2[i];
    for (int i = threadIdx.x; i < WOUT * 8; i += blockDim.x) sw3[i] = w3[i];
    for (int i = threadIdx.x; i < WOUT; i += blockDim.x) sb3[i] = b3[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gwarp = (long long)blockIdx.x * (WG_THREADS / 32) + warp;
    const long long nwarps = (long long)gridDim.x * (WG_THREADS / 32);
    float (*myrec)[WG_REC] = rec[warp];

    // parameter ownership: lanes 0..W-1 -> row `lane` of dW3 (G = g3[lane], V = h2); lanes 24..31 own nothing when W = 24;
    // lanes W..W+7 -> dW2 row (G = g2, V = h1); the next 8 lanes -> dW1 row (G = g1, V = x padded to 8)
    int g_off, v_off;
    bool owner = true;
    if (lane < WOUT) { g_off = 40 + lane; v_off = 16; }
    else if (lane < WOUT + 8) { g_off = 32 + (lane - WOUT); v_off = 8; }
    else if (lane < WOUT + 16) { g_off = 24 + (lane - WOUT - 8); v_off = 0; }
    else { g_off = 0; v_off = 0; owner = false; }
    float acc[8], accb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;

    for (long long base = gwarp * 32; base < rows; base += nwarps * 32) {
        const long long r = base + lane;
        float *me = myrec[lane];
        if (r < rows) {
            const float *x = in + r * in_stride;
            const float x0 = x[0], x1 = x[1], x2 = x[2];
            float h1[8], h2[8], g3[WOUT], g2[8], g1[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) h1[o] = fmaxf(sb1[o] + sw1[o * 3 + 0] * x0 + sw1[o * 3 + 1] * x1 + sw1[o * 3 + 2] * x2, 0.f);
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                float a = sb2[o];
#pragma unroll
                for (int i = 0; i < 8; ++i) a += sw2[o * 8 + i] * h1[i];
                h2[o] = fmaxf(a, 0.f);
            }
            const float *go = g_out + r * WOUT;
#pragma unroll
            for (int o4 = 0; o4 < WOUT; o4 += 4) {
                const float4 gv = __ldg(reinterpret_cast<const float4 *>(go + o4));
                const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float a = sb3[o4 + u];
#pragma unroll
                    for (int i = 0; i < 8; ++i) a += sw3[(o4 + u) * 8 + i] * h2[i];
                    g3[o4 + u] = a > 0.f ? gg[u] : 0.f;           // ReLU mask of the output layer (threshold_backward)
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float a = 0.f;
#pragma unroll
                for (int o = 0; o < WOUT; ++o) a += sw3[o * 8 + i] * g3[o];
                g2[i] = h2[i] > 0.f ? a : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float a = 0.f;
#pragma unroll
                for (int o = 0; o < 8; ++o) a += sw2[o * 8 + i] * g2[o];
                g1[i] = h1[i] > 0.f ? a : 0.f;
            }
            if (g_in != nullptr) {
                float gx[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int o = 0; o < 8; ++o) { gx[0] += sw1[o * 3 + 0] * g1[o]; gx[1] += sw1[o * 3 + 1] * g1[o]; gx[2] += sw1[o * 3 + 2] * g1[o]; }
                g_in[r * 3 + 0] = gx[0]; g_in[r * 3 + 1] = gx[1]; g_in[r * 3 + 2] = gx[2];
            }
            me[0] = x0; me[1] = x1; me[2] = x2; me[3] = 0.f; me[4] = 0.f; me[5] = 0.f; me[6] = 0.f; me[7] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { me[8 + i] = h1[i]; me[16 + i] = h2[i]; me[24 + i] = g1[i]; me[32 + i] = g2[i]; }
#pragma unroll
            for (int o = 0; o < WOUT; ++o) me[40 + o] = g3[o];
        } else {
#pragma unroll
            for (int i = 0; i < 40 + WOUT; ++i) me[i] = 0.f;     // a padded row contributes nothing
        }
        __syncwarp();
        if (owner) {
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
                const float *q = myrec[rr];
                const float gq = q[g_off];
                const float4 va = *reinterpret_cast<const float4 *>(q + v_off), vb = *reinterpret_cast<const float4 *>(q + v_off + 4);
                acc[0] += gq * va.x; acc[1] += gq * va.y; acc[2] += gq * va.z; acc[3] += gq * va.w;
                acc[4] += gq * vb.x; acc[5] += gq * vb.y; acc[6] += gq * vb.z; acc[7] += gq * vb.w;
                accb += gq;
            }
        }
        __syncwarp();
    }
    // per-warp partials: [gwarp][lane][9]
    float *p = partial + (gwarp * 32 + lane) * 9;
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = acc[j];
    p[8] = accb;
}

// sums the per-warp partials in warp order and scatters them to the six gradient tensors
template <int WOUT>
__global__ void __launch_bounds__(32 * 9)
weightnet_grad_reduce_kernel(long long nwarps, const float *__restrict__ partial, float *__restrict__ gw1, float *__restrict__ gb1,
                             float *__restrict__ gw2, float *__restrict__ gb2, float *__restrict__ gw3, float *__restrict__ gb3) {
    const int lane = threadIdx.x / 9, j = threadIdx.x - lane * 9;
    float a = 0.f;
    for (long long w = 0; w < nwarps; ++w) a += partial[(w * 32 + lane) * 9 + j];
    if (lane < WOUT) { if (j < 8) gw3[lane * 8 + j] = a; else gb3[lane] = a; }
    else if (lane < WOUT + 8) { const int o = lane - WOUT; if (j < 8) gw2[o * 8 + j] = a; else gb2[o] = a; }
    else if (lane < WOUT + 16) { const int o = lane - WOUT - 8; if (j < 3) gw1[o * 3 + j] = a; else if (j == 8) gb1[o] = a; }
}

static inline long long wg_grid(long long rows) {
    long long ctas = (rows + WG_THREADS - 1) / WG_THREADS;
    const long long cap = 2LL * device_sms();      // the final reduction walks one partial per warp
    return ctas < cap ? (ctas < 1 ? 1 : ctas) : cap;
}

}  // namespace kdpc

using namespace kdpc;

KDPC_API long long kdpc_weightnet_grad_ws_bytes(long long rows) {
    return wg_grid(rows) * (WG_THREADS / 32) * 32 * 9 * (long long)sizeof(float);
}

/* Gradients of WeightNet(3 -> 8 -> 8 -> wout), wout in {8, 16}: in [rows, in_stride] (first 3 columns = localized xyz),
 * g_out [rows, wout]; writes gw1 [8,3] gb1 [8] gw2 [8,8] gb2 [8] gw3 [wout,8] gb3 [wout] and, when g_in != NULL, g_in [rows,3]. */
KDPC_API int kdpc_weightnet_grad(long long rows, const float *in, int in_stride, int wout, const float *g_out,
                                 const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                                 const float *b3, void *ws, float *gw1, float *gb1, float *gw2, float *gb2, float *gw3,
                                 float *gb3, float *g_in, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(rows > 0 && in && g_out && w1 && b1 && w2 && b2 && w3 && b3 && ws && gw1 && gb1 && gw2 && gb2 && gw3 && gb3);
    KDPC_CHECK_ARGS(in_stride >= 3);
    if ((reinterpret_cast<uintptr_t>(g_out) % 16) != 0) return KDPC_EINVAL;
    const long long grid = wg_grid(rows);
    const long long nwarps = grid * (WG_THREADS / 32);
    cudaStream_t st = to_stream(stream);
    float *partial = reinterpret_cast<float *>(ws);
    if (wout == 16) {
        weightnet_grad_kernel<16><<<(unsigned)grid, WG_THREADS, 0, st>>>(rows, in, in_stride, g_out, w1, b1, w2, b2, w3, b3, partial, g_in);
        weightnet_grad_reduce_kernel<16><<<1, 32 * 9, 0, st>>>(nwarps, partial, gw1, gb1, gw2, gb2, gw3, gb3);
    } else if (wout == 8) {
        weightnet_grad_kernel<8><<<(unsigned)grid, WG_THREADS, 0, st>>>(rows, in, in_stride, g_out, w1, b1, w2, b2, w3, b3, partial, g_in);
        weightnet_grad_reduce_kernel<8><<<1, 32 * 9, 0, st>>>(nwarps, partial, gw1, gb1, gw2, gb2, gw3, gb3);
    } else {
        return KDPC_EUNSUPPORTED;
    }
    KDPC_RETURN_LAST();
}
