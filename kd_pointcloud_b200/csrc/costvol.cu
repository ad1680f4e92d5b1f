// Bidirectional cost-volume pieces (reference CrossLayerLight.cross, pointconv_util.py:1826-1850).
//
//  * costvol_pre : gather p2[idx] + broadcast p1 + pos(xyz2[idx]-xyz1) + activation in ONE pass,
//                  writing [B,S,K,D] point-major.  The reference builds this from two grouping
//                  kernels, a `repeat` copy of p1, a cuDNN 3->D conv and three elementwise passes.
//  * max_over_k  : the F.max_pool2d over the neighbour axis (pointconv_util.py:1848), with arg-max
//                  kept for the backward.
#include "common.cuh"

namespace kdpc {

__device__ __forceinline__ float act_leaky(float v, float slope) { return v > 0.f ? v : v * slope; }

__global__ void __launch_bounds__(256)
costvol_pre_kernel(long long total, int s, int n, int k, int dvec, const float *__restrict__ xyz1,
                   const float *__restrict__ xyz2, const float *__restrict__ p1, const float *__restrict__ p2,
                   const int *__restrict__ idx, const float *__restrict__ pos_w, const float *__restrict__ pos_b,
                   float slope, float *__restrict__ out) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long r = e / dvec;                         // r = (b*s + i)*k + kk
    const int dv = (int)(e - r * dvec);
    const long long bi = r / k;                           // b*s + i
    const long long b = bi / s;
    const int src = __ldg(idx + r);
    const float *q = xyz1 + bi * 3;
    const float *c = xyz2 + ((size_t)b * n + src) * 3;
    const float dx = c[0] - q[0], dy = c[1] - q[1], dz = c[2] - q[2];
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p2) + ((size_t)b * n + src) * dvec + dv);
    const float4 p = __ldg(reinterpret_cast<const float4 *>(p1) + (size_t)bi * dvec + dv);
    const float4 pb = __ldg(reinterpret_cast<const float4 *>(pos_b) + dv);
    const float *pw = pos_w + (size_t)dv * 12;            // 4 output channels x 3
    float4 o;
    o.x = act_leaky((a.x + p.x) + (pb.x + pw[0] * dx + pw[1] * dy + pw[2] * dz), slope);
    o.y = act_leaky((a.y + p.y) + (pb.y + pw[3] * dx + pw[4] * dy + pw[5] * dz), slope);
    o.z = act_leaky((a.z + p.z) + (pb.z + pw[6] * dx + pw[7] * dy + pw[8] * dz), slope);
    o.w = act_leaky((a.w + p.w) + (pb.w + pw[9] * dx + pw[10] * dy + pw[11] * dz), slope);
    st_stream_f4(reinterpret_cast<float4 *>(out) + e, o);
}

__global__ void __launch_bounds__(256)
max_over_k_kernel(long long total, int k, int dvec, const float *__restrict__ in, float *__restrict__ out,
                  int *__restrict__ arg) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long r = e / dvec;
    const int dv = (int)(e - r * dvec);
    const float4 *ip = reinterpret_cast<const float4 *>(in) + (size_t)r * k * dvec + dv;
    float4 m = ld_stream_f4(ip);
    int4 am = make_int4(0, 0, 0, 0);
    for (int kk = 1; kk < k; ++kk) {
        const float4 v = ld_stream_f4(ip + (size_t)kk * dvec);
        if (v.x > m.x) { m.x = v.x; am.x = kk; }
        if (v.y > m.y) { m.y = v.y; am.y = kk; }
        if (v.z > m.z) { m.z = v.z; am.z = kk; }
        if (v.w > m.w) { m.w = v.w; am.w = kk; }
    }
    reinterpret_cast<float4 *>(out)[e] = m;
    if (arg != nullptr) reinterpret_cast<int4 *>(arg)[e] = am;
}

}  // namespace kdpc

using namespace kdpc;

KDPC_API int kdpc_costvol_pre(int b, int s, int n, int k, int d, const float *xyz1, const float *xyz2,
                              const float *p1, const float *p2, const int *idx, const float *pos_w,
                              const float *pos_b, float slope, float *out, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(xyz1 && xyz2 && p1 && p2 && idx && pos_w && pos_b && out && b > 0 && s > 0 && n > 0 && k > 0 && d > 0);
    if (d % 4 != 0) return KDPC_EUNSUPPORTED;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2) |
                         reinterpret_cast<uintptr_t>(pos_b) | reinterpret_cast<uintptr_t>(out);
    if (al % 16 != 0) return KDPC_EINVAL;
    const long long total = (long long)b * s * k * (d / 4);
    costvol_pre_kernel<<<(unsigned)div_up_ll(total, 256), 256, 0, to_stream(stream)>>>(
        total, s, n, k, d / 4, xyz1, xyz2, p1, p2, idx, pos_w, pos_b, slope, out);
    KDPC_RETURN_LAST();
}

KDPC_API int kdpc_max_over_k(long long rows, int k, int d, const float *in, float *out, int *arg,
                             kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(in && out && rows > 0 && k > 0 && d > 0);
    if (d % 4 != 0) return KDPC_EUNSUPPORTED;
    const uintptr_t al = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(arg);
    if (al % 16 != 0) return KDPC_EINVAL;
    const long long total = rows * (d / 4);
    max_over_k_kernel<<<(unsigned)div_up_ll(total, 256), 256, 0, to_stream(stream)>>>(total, k, d / 4, in, out, arg);
    KDPC_RETURN_LAST();
}
