// Three-neighbour interpolation kernels for sm_100a.
//
//  * three_interpolate (channel-major pointnet2 API): replaces three_interpolate_kernel_fast
//    (reference pointnet2/src/interpolate_gpu.cu:77-97); same fma order as the reference's SASS.
//  * interp3 (point-major): the inverse-distance interpolation shared by UpsampleFlow and
//    PointWarping (pointconv_util.py:2131-2139, 2164-2171) fused into one pass: weights from the
//    coordinates + weighted sum of three feature rows, one 128-bit access per thread.
#include "common.cuh"

namespace kdpc {

// thread = (point i, group of TI_CG channels): the 3 indices / weights are read once per channel group (L2 hits),
// TI_CG independent gather chains per thread and c / TI_CG times more threads than one-thread-per-point
// (which left 256 CTAs of serial 64-channel loops: 1.7x slower than the reference's thread-per-element kernel).
constexpr int TI_CG = 4;
__global__ void __launch_bounds__(256)
three_interpolate_cm_kernel(int c, int m, int n, const float *__restrict__ f, const int *__restrict__ idx,
                            const float *__restrict__ w, float *__restrict__ out) {
    const int b = blockIdx.z, c0 = blockIdx.y * TI_CG;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int *ip = idx + ((size_t)b * n + i) * 3;
    const float *wp = w + ((size_t)b * n + i) * 3;
    const int i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
    const float w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
    const float *fb = f + ((size_t)b * c + c0) * m;
    float *ob = out + ((size_t)b * c + c0) * n;
#pragma unroll
    for (int ci = 0; ci < TI_CG; ++ci) {
        if (c0 + ci < c) {
            const float *fr = fb + (size_t)ci * m;
            // w0*p0 + w1*p1 + w2*p2 as nvcc contracts it in the reference: fma(w2,p2, fma(w0,p0, rn(w1*p1)))
            float t = __fmul_rn(w1, __ldg(fr + i1));
            t = __fmaf_rn(w0, __ldg(fr + i0), t);
            ob[(size_t)ci * n + i] = __fmaf_rn(w2, __ldg(fr + i2), t);
        }
    }
}

__device__ __forceinline__ void interp3_weights(const float *__restrict__ q, const float *__restrict__ cb,
                                                int i0, int i1, int i2, float &w0, float &w1, float &w2) {
    const float qx = q[0], qy = q[1], qz = q[2];
    float r[3];
    const int ii[3] = {i0, i1, i2};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float *cp = cb + (size_t)ii[j] * 3;
        const float dx = cp[0] - qx, dy = cp[1] - qy, dz = cp[2] - qz;
        // torch.norm(dim=3).clamp(min=1e-10); 1.0 / dist          (pointconv_util.py:2133-2135)
        const float dist = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-10f);
        r[j] = 1.0f / dist;
    }
    const float norm = (r[0] + r[1]) + r[2];
    w0 = r[0] / norm; w1 = r[1] / norm; w2 = r[2] / norm;
}

// thread = one 16-byte (VEC = 4) or 4-byte (VEC = 1) piece of an output row; the cvec pieces of a point are consecutive
// threads.  The indices and the three inverse-distance weights of a point (3 square roots, 6 divisions, 15 loads) are
// computed ONCE per point and warp - by the lane that holds the point's first piece, or by lane 0 when the point began in
// the previous warp - and handed to the point's other lanes by shuffle (every lane used to recompute them: 16 times per
// point at C = 64, which made the kernel instruction-bound at 19 % of the copy peak).  IT: 32-bit index arithmetic
// whenever the piece count allows (two 64-bit divisions per thread otherwise).  Same values, bit for bit.
template <int VEC, typename IT>
__global__ void __launch_bounds__(256)
interp3_kernel(IT total, int n, int s, int cvec, const float *__restrict__ q_xyz, const float *__restrict__ c_xyz,
               const int *__restrict__ idx, const float *__restrict__ feat, float *__restrict__ out,
               float *__restrict__ w_out) {
    IT e = (IT)blockIdx.x * (IT)256 + (IT)threadIdx.x;
    const bool live = e < total;
    if (!live) e = total - 1;                             // the lane stays for the warp-wide exchange; it stores nothing
    const IT r = e / (IT)cvec;                            // r = b*n + i
    const int cv = (int)(e - r * (IT)cvec);
    const IT b = r / (IT)n;
    const int lane = threadIdx.x & 31;
    const int leader = max(lane - cv, 0);                 // lane of this warp that computes this point's weights
    int i0 = 0, i1 = 0, i2 = 0;
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
    if (cv == 0 || lane == 0) {
        const int *ip = idx + (size_t)r * 3;
        i0 = ip[0]; i1 = ip[1]; i2 = ip[2];
        interp3_weights(q_xyz + (size_t)r * 3, c_xyz + (size_t)b * s * 3, i0, i1, i2, w0, w1, w2);
        if (w_out != nullptr && cv == 0 && live) {
            w_out[(size_t)r * 3 + 0] = w0; w_out[(size_t)r * 3 + 1] = w1; w_out[(size_t)r * 3 + 2] = w2;
        }
    }
    i0 = __shfl_sync(0xffffffffu, i0, leader);
    i1 = __shfl_sync(0xffffffffu, i1, leader);
    i2 = __shfl_sync(0xffffffffu, i2, leader);
    w0 = __shfl_sync(0xffffffffu, w0, leader);
    w1 = __shfl_sync(0xffffffffu, w1, leader);
    w2 = __shfl_sync(0xffffffffu, w2, leader);
    if (!live) return;
    const size_t fb = (size_t)b * s * cvec;
    if (VEC == 4) {
        const float4 *f4 = reinterpret_cast<const float4 *>(feat);
        const float4 a = __ldg(f4 + fb + (size_t)i0 * cvec + cv);
        const float4 bq = __ldg(f4 + fb + (size_t)i1 * cvec + cv);
        const float4 cq = __ldg(f4 + fb + (size_t)i2 * cvec + cv);
        float4 o;
        o.x = (w0 * a.x + w1 * bq.x) + w2 * cq.x;
        o.y = (w0 * a.y + w1 * bq.y) + w2 * cq.y;
        o.z = (w0 * a.z + w1 * bq.z) + w2 * cq.z;
        o.w = (w0 * a.w + w1 * bq.w) + w2 * cq.w;
        reinterpret_cast<float4 *>(out)[e] = o;
    } else {
        const float a = __ldg(feat + fb + (size_t)i0 * cvec + cv);
        const float bq = __ldg(feat + fb + (size_t)i1 * cvec + cv);
        const float cq = __ldg(feat + fb + (size_t)i2 * cvec + cv);
        out[e] = (w0 * a + w1 * bq) + w2 * cq;
    }
}

}  // namespace kdpc

using namespace kdpc;

KDPC_API int kdpc_three_interpolate(int b, int c, int m, int n, const float *f, const int *idx, const float *w,
                                    float *out, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(f && idx && w && out && b > 0 && c > 0 && m > 0 && n > 0);
    if (b > 65535 || c > 65535 * TI_CG) return KDPC_EUNSUPPORTED;
    dim3 grid((n + 255) / 256, (c + TI_CG - 1) / TI_CG, b);
    three_interpolate_cm_kernel<<<grid, 256, 0, to_stream(stream)>>>(c, m, n, f, idx, w, out);
    KDPC_RETURN_LAST();
}

KDPC_API int kdpc_interp3(int b, int n, int s, int c, const float *q_xyz, const float *c_xyz, const int *idx,
                          const float *feat, float *out, float *w_out, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(q_xyz && c_xyz && idx && feat && out && b > 0 && n > 0 && s > 0 && c > 0);
    cudaStream_t st = to_stream(stream);
    const bool vec = (c % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
    const int cvec = vec ? c / 4 : c;
    const long long total = (long long)b * n * cvec;
    const unsigned grid = (unsigned)div_up_ll(total, 256);
    if (total < (1ll << 31)) {                               // 32-bit index arithmetic (every shape of the model)
        if (vec) interp3_kernel<4, unsigned><<<grid, 256, 0, st>>>((unsigned)total, n, s, cvec, q_xyz, c_xyz, idx, feat, out, w_out);
        else interp3_kernel<1, unsigned><<<grid, 256, 0, st>>>((unsigned)total, n, s, cvec, q_xyz, c_xyz, idx, feat, out, w_out);
    } else {
        if (div_up_ll(total, 256) >= (1ll << 31)) return KDPC_EUNSUPPORTED;
        if (vec) interp3_kernel<4, long long><<<grid, 256, 0, st>>>(total, n, s, cvec, q_xyz, c_xyz, idx, feat, out, w_out);
        else interp3_kernel<1, long long><<<grid, 256, 0, st>>>(total, n, s, cvec, q_xyz, c_xyz, idx, feat, out, w_out);
    }
    KDPC_RETURN_LAST();
}
