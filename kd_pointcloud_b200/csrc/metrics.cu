// On-device scene-flow evaluation metrics, sm_100a.
//
// Reference: evaluate_3d / evaluate_2d (evaluation_utils.py:17-50) on numpy arrays, fed by
// evaluate_bid_pointconv.py:128-145 after four .cpu().numpy() copies per batch, and
// geometry.get_batch_2d_flow / project_3d_to_2d (utils/geometry.py:6-65).  Here ONE launch reads the
// prediction, the ground truth and the first cloud where they already are and leaves the six batch means in device
// memory: no synchronisation, no [B,N,3] traffic over PCIe.
//   * per point, every float32 operation of the reference in the reference's order with IEEE rounding
//     (__fmul_rn / __fadd_rn / __fsqrt_rn / __fdiv_rn: no FMA contraction), so the threshold tests are
//     bit-identical to numpy's and the four accuracy COUNTS are exact integers;
//   * the two error sums are float64 (numpy sums float32 pairwise; both agree with the exact sum to ~1e-7 relative);
//   * deterministic: fixed grid and strides, shared-memory tree per block, the last block adds the block partials in
//     index order (no floating-point atomics).
#include "common.cuh"

namespace kdpc {

constexpr int MET_THREADS = 256;
constexpr int MET_BLOCKS = 256;        // workspace capacity (partials); the grid is min(SM count, MET_BLOCKS)
static inline int met_grid() { int n = num_sms(); return n < MET_BLOCKS ? n : MET_BLOCKS; }

struct MetricPartial {
    double l2, epe2d;
    unsigned long long strict, relax, outlier, acc2d;
};

// x = (X f + cx Z + constx) / (Z + constz), utils/geometry.py:61-65 (float32 throughout)
__device__ __forceinline__ void project(float X, float Y, float Z, const float *c, float &px, float &py) {
    const float den = __fadd_rn(Z, c[5]);
    px = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, c[0]), __fmul_rn(c[1], Z)), c[3]), den);
    py = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(Y, c[0]), __fmul_rn(c[2], Z)), c[4]), den);
}
__device__ __forceinline__ float norm3(float x, float y, float z) {          // np.linalg.norm(axis=-1): sqrt((x^2 + y^2) + z^2)
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}
__device__ __forceinline__ float norm2(float x, float y) { return __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))); }

__global__ void __launch_bounds__(MET_THREADS)
flow_metrics_kernel(int b, int n, int point_major, const float *__restrict__ pred, const float *__restrict__ gt,
                    const float *__restrict__ pc1, const float *__restrict__ calib, MetricPartial *partials,
                    unsigned *counter, float *out) {
    __shared__ MetricPartial red[MET_THREADS / 32];
    __shared__ bool last;
    MetricPartial acc = {0.0, 0.0, 0ull, 0ull, 0ull, 0ull};
    const float dflt[6] = {-1050.f, 479.5f, 269.5f, 0.f, 0.f, 0.f};           // project_3d_to_2d defaults (FlyingThings3D)
    const long long items = (long long)b * n;
    for (long long e = (long long)blockIdx.x * MET_THREADS + threadIdx.x; e < items; e += (long long)gridDim.x * MET_THREADS) {
        const int bi = (int)(e / n), p = (int)(e - (long long)bi * n);
        const int sc = point_major ? 1 : n;
        const float *pp = pred + (size_t)bi * 3 * n + (point_major ? (size_t)p * 3 : (size_t)p);
        const float fx = pp[0], fy = pp[sc], fz = pp[2 * sc];
        const float *gp = gt + (size_t)e * 3;
        const float gx = gp[0], gy = gp[1], gz = gp[2];
        // evaluate_3d, evaluation_utils.py:22-33
        const float l2 = norm3(__fsub_rn(gx, fx), __fsub_rn(gy, fy), __fsub_rn(gz, fz));
        const float rel = __fdiv_rn(l2, __fadd_rn(norm3(gx, gy, gz), 1e-4f));
        acc.l2 += (double)l2;
        acc.strict += (l2 < 0.05f || rel < 0.05f) ? 1ull : 0ull;
        acc.relax += (l2 < 0.1f || rel < 0.1f) ? 1ull : 0ull;
        acc.outlier += (l2 > 0.3f || rel > 0.1f) ? 1ull : 0ull;
        if (pc1 != nullptr) {
            // get_batch_2d_flow(pc1, pc1 + gt, pc1 + pred), evaluate_bid_pointconv.py:138-141 + geometry.py:41-58
            const float *c = calib != nullptr ? calib + (size_t)bi * 6 : dflt;
            const float *xp = pc1 + (size_t)e * 3;
            const float x = xp[0], y = xp[1], z = xp[2];
            float px1, py1, px2, py2, pxg, pyg;
            project(x, y, z, c, px1, py1);
            project(__fadd_rn(x, fx), __fadd_rn(y, fy), __fadd_rn(z, fz), c, px2, py2);
            project(__fadd_rn(x, gx), __fadd_rn(y, gy), __fadd_rn(z, gz), c, pxg, pyg);
            const float fpx = __fsub_rn(px2, px1), fpy = __fsub_rn(py2, py1);
            const float fgx = __fsub_rn(pxg, px1), fgy = __fsub_rn(pyg, py1);
            // evaluate_2d, evaluation_utils.py:42-48
            const float epe = norm2(__fsub_rn(fgx, fpx), __fsub_rn(fgy, fpy));
            const float rel2 = __fdiv_rn(epe, __fadd_rn(norm2(fgx, fgy), 1e-5f));
            acc.epe2d += (double)epe;
            acc.acc2d += (epe < 3.f || rel2 < 0.05f) ? 1ull : 0ull;
        }
    }
    // block reduction (fixed tree), then the last block adds the block partials in index order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.l2 += __shfl_xor_sync(0xffffffffu, acc.l2, o);
        acc.epe2d += __shfl_xor_sync(0xffffffffu, acc.epe2d, o);
        acc.strict += __shfl_xor_sync(0xffffffffu, acc.strict, o);
        acc.relax += __shfl_xor_sync(0xffffffffu, acc.relax, o);
        acc.outlier += __shfl_xor_sync(0xffffffffu, acc.outlier, o);
        acc.acc2d += __shfl_xor_sync(0xffffffffu, acc.acc2d, o);
    }
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        MetricPartial t = red[0];
        for (int w = 1; w < MET_THREADS / 32; ++w) {
            t.l2 += red[w].l2; t.epe2d += red[w].epe2d;
            t.strict += red[w].strict; t.relax += red[w].relax; t.outlier += red[w].outlier; t.acc2d += red[w].acc2d;
        }
        partials[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        MetricPartial t = {0.0, 0.0, 0ull, 0ull, 0ull, 0ull};
        const volatile MetricPartial *vp = partials;
        for (unsigned i = 0; i < gridDim.x; ++i) {
            t.l2 += vp[i].l2; t.epe2d += vp[i].epe2d;
            t.strict += vp[i].strict; t.relax += vp[i].relax; t.outlier += vp[i].outlier; t.acc2d += vp[i].acc2d;
        }
        const double inv = 1.0 / (double)items;
        out[0] = (float)(t.l2 * inv);                 // EPE3D
        out[1] = (float)((double)t.strict * inv);     // Acc3DS
        out[2] = (float)((double)t.relax * inv);      // Acc3DR
        out[3] = (float)((double)t.outlier * inv);    // Outliers3D
        out[4] = (float)(t.epe2d * inv);              // EPE2D
        out[5] = (float)((double)t.acc2d * inv);      // Acc2D
        *counter = 0u;                                // ready for the next launch
    }
}

}  // namespace kdpc

using namespace kdpc;

KDPC_API long long kdpc_flow_metrics_workspace_bytes(void) { return (long long)MET_BLOCKS * sizeof(MetricPartial) + 16; }

KDPC_API int kdpc_flow_metrics(int b, int n, int point_major, const float *pred, const float *gt, const float *pc1,
                               const float *calib, void *ws, float *out, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(pred && gt && ws && out && b > 0 && n > 0);
    if ((reinterpret_cast<uintptr_t>(ws) % 8) != 0) return KDPC_EINVAL;
    MetricPartial *partials = reinterpret_cast<MetricPartial *>(ws);
    unsigned *counter = reinterpret_cast<unsigned *>(partials + MET_BLOCKS);
    flow_metrics_kernel<<<met_grid(), MET_THREADS, 0, to_stream(stream)>>>(b, n, point_major ? 1 : 0, pred, gt, pc1, calib,
                                                                         partials, counter, out);
    KDPC_RETURN_LAST();
}
