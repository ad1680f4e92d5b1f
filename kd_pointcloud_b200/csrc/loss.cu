// Fused multi-scale flow loss + distillation hint loss (forward AND gradient in one pass), sm_100a.
//
// Reference: multiScaleLoss (loss_functions.py:6-25; copies models_bid_pointconv.py:545-563) builds the
// ground-truth pyramid with three chained index_points_gather calls (permute + contiguous + gather kernel +
// permute each), then per scale permute / subtract / norm / sum / mean / scale / add: ~20 launches, and the
// distillation losses (loss_functions.py:27-36, 83-96, 201-219) call it twice (vs. the teacher's flow and vs.
// the ground truth) and add 0.5 * sum((f_student - f_teacher)^2) hint terms.  Here:
//   * flow_loss_kernel: ONE launch covers every (scale, batch, point) of every target.  The pyramid is never
//     materialised: the FPS index chain is composed on the fly (gt_i[b,n] = gt[b, idx_0[idx_1[..[n]]]]).
//     The same pass writes d loss / d pred (the norm's sub-gradient is 0 at 0, as torch.norm's).
//   * hint_loss_kernel: 0.5 * w * sum (a - b)^2 and its gradient w (a - b).
// Reductions are deterministic: fixed grid, fixed per-thread strides, a shared-memory tree per block, and the
// LAST block to finish adds the block partials in index order (no floating-point atomics).
#include "common.cuh"

namespace kdpc {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_BLOCKS = 256;       // workspace capacity (partials); the grid is min(SM count, LOSS_BLOCKS)
static inline int loss_grid() { int n = num_sms(); return n < LOSS_BLOCKS ? n : LOSS_BLOCKS; }
constexpr int LOSS_MAX_SCALES = 4;
constexpr int LOSS_MAX_TARGETS = 2;

struct FlowLossArgs {
    int b, nscales, ntargets, n0;
    int point_major;                              // pred / grad layout: 0 = [B,3,N_i], 1 = [B,N_i,3]
    int n[LOSS_MAX_SCALES];                       // points per scale (finest first)
    const float *pred[LOSS_MAX_SCALES];           // [B,3,N_i] channel-major (the model's output layout)
    float *grad[LOSS_MAX_SCALES];                 // same layout, or nullptr
    const int *fps[LOSS_MAX_SCALES];              // fps[i]: int32 [B,N_{i+1}] -> indices into scale i (i < nscales-1)
    float alpha[LOSS_MAX_SCALES];
    const float *target[LOSS_MAX_TARGETS];        // [B,N0,3] point-major
    float weight[LOSS_MAX_TARGETS];               // includes 1/B (the mean over the batch)
};

__device__ __forceinline__ float block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (warp == 0) {
        t = lane < (LOSS_THREADS / 32) ? red[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;                                      // valid on thread 0
}

// partials[gridDim.x] + counter: the last block sums the partials in index order and adds to *loss.
__device__ __forceinline__ void finish(float block_total, float *partials, unsigned *counter, float *loss) {
    __shared__ bool last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = block_total;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        float t = 0.f;
        for (unsigned i = 0; i < gridDim.x; ++i) t += reinterpret_cast<volatile float *>(partials)[i];
        *loss += t;                                // launches on one stream are ordered: plain read-modify-write
        *counter = 0u;                             // ready for the next launch
    }
}

__global__ void __launch_bounds__(LOSS_THREADS)
flow_loss_kernel(const FlowLossArgs a, float *partials, unsigned *counter, float *loss) {
    __shared__ float red[LOSS_THREADS / 32];
    float acc = 0.f;
    for (int s = 0; s < a.nscales; ++s) {
        const int ns = a.n[s];
        const long long items = (long long)a.b * ns;
        for (long long e = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x; e < items;
             e += (long long)gridDim.x * LOSS_THREADS) {
            const int bi = (int)(e / ns), p = (int)(e - (long long)bi * ns);
            int j = p;                             // compose the FPS chain down to the finest level
            for (int l = s - 1; l >= 0; --l) j = __ldg(a.fps[l] + (size_t)bi * a.n[l + 1] + j);
            const int sc = a.point_major ? 1 : ns;                           // channel stride
            const size_t po = (size_t)bi * 3 * ns + (a.point_major ? (size_t)p * 3 : (size_t)p);
            const float *pp = a.pred[s] + po;
            const float px = pp[0], py = pp[sc], pz = pp[2 * sc];
            float gx = 0.f, gy = 0.f, gz = 0.f;
            for (int t = 0; t < a.ntargets; ++t) {
                const float *tp = a.target[t] + ((size_t)bi * a.n0 + j) * 3;
                const float dx = px - tp[0], dy = py - tp[1], dz = pz - tp[2];
                const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
                const float w = a.alpha[s] * a.weight[t];
                acc += w * nrm;
                const float g = nrm > 0.f ? w / nrm : 0.f;
                gx += g * dx; gy += g * dy; gz += g * dz;
            }
            if (a.grad[s] != nullptr) {
                float *gp = a.grad[s] + po;
                gp[0] = gx; gp[sc] = gy; gp[2 * sc] = gz;
            }
        }
    }
    finish(block_sum(acc, red), partials, counter, loss);
}

__global__ void __launch_bounds__(LOSS_THREADS)
hint_loss_kernel(long long n, const float *__restrict__ fs, const float *__restrict__ ft, float weight,
                 float *__restrict__ grad, float *partials, unsigned *counter, float *loss) {
    __shared__ float red[LOSS_THREADS / 32];
    float acc = 0.f;
    for (long long e = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x; e < n; e += (long long)gridDim.x * LOSS_THREADS) {
        const float d = fs[e] - __ldg(ft + e);
        acc += d * d;
        if (grad != nullptr) grad[e] = weight * d;
    }
    finish(0.5f * weight * block_sum(acc, red), partials, counter, loss);
}

}  // namespace kdpc

using namespace kdpc;

KDPC_API long long kdpc_loss_workspace_bytes(void) { return (long long)(LOSS_BLOCKS + 1) * 4; }

KDPC_API int kdpc_flow_loss(int b, int nscales, int point_major, const int *n, const float *const *pred, float *const *grad_pred,
                            const int *const *fps_idx, const float *alpha, int ntargets, const float *const *target,
                            const float *weight, void *ws, float *loss, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(n && pred && alpha && target && weight && ws && loss && b > 0);
    if (nscales < 1 || nscales > LOSS_MAX_SCALES || ntargets < 1 || ntargets > LOSS_MAX_TARGETS) return KDPC_EUNSUPPORTED;
    FlowLossArgs a;
    a.b = b; a.nscales = nscales; a.ntargets = ntargets; a.n0 = n[0]; a.point_major = point_major ? 1 : 0;
    for (int s = 0; s < LOSS_MAX_SCALES; ++s) {
        const bool v = s < nscales;
        a.n[s] = v ? n[s] : 0;
        a.pred[s] = v ? pred[s] : nullptr;
        a.grad[s] = (v && grad_pred) ? grad_pred[s] : nullptr;
        a.fps[s] = (s + 1 < nscales && fps_idx) ? fps_idx[s] : nullptr;
        a.alpha[s] = v ? alpha[s] : 0.f;
        if (v && (a.n[s] <= 0 || !a.pred[s])) return KDPC_EINVAL;
        if (s + 1 < nscales && !a.fps[s]) return KDPC_EINVAL;
    }
    for (int t = 0; t < LOSS_MAX_TARGETS; ++t) {
        a.target[t] = t < ntargets ? target[t] : nullptr;
        a.weight[t] = t < ntargets ? weight[t] : 0.f;
        if (t < ntargets && !a.target[t]) return KDPC_EINVAL;
    }
    float *partials = reinterpret_cast<float *>(ws);
    unsigned *counter = reinterpret_cast<unsigned *>(ws) + LOSS_BLOCKS;
    flow_loss_kernel<<<loss_grid(), LOSS_THREADS, 0, to_stream(stream)>>>(a, partials, counter, loss);
    KDPC_RETURN_LAST();
}

KDPC_API int kdpc_hint_loss(long long n, const float *fs, const float *ft, float weight, float *grad_fs, void *ws,
                            float *loss, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(fs && ft && ws && loss && n > 0);
    float *partials = reinterpret_cast<float *>(ws);
    unsigned *counter = reinterpret_cast<unsigned *>(ws) + LOSS_BLOCKS;
    hint_loss_kernel<<<loss_grid(), LOSS_THREADS, 0, to_stream(stream)>>>(n, fs, ft, weight, grad_fs, partials, counter, loss);
    KDPC_RETURN_LAST();
}
