// Adam over ALL parameter tensors of the student in one launch per 96 tensors (distilTrain.py:134-135 + optimizer.step(),
// :182: torch.optim.Adam(lr, betas, eps, weight_decay)).  torch's capturable foreach implementation walks the 226
// parameter tensors with a dozen multi-tensor element-wise launches (1.9 ms of a 28 ms KD step for 8 M parameters =
// 220 MB of traffic); here the pointer table travels BY VALUE in the kernel parameters (so a CUDA graph records it and
// eager steps may pass new .grad addresses every time), a CTA owns one 4096-element chunk of one tensor, and the update
// is a single pass: read p, g, m, v - write p, m, v.  Arithmetic = torch.optim.Adam's (default path) in fp32:
//     g += wd p;  m = lerp(m, g, 1 - b1);  v = b2 v + (1 - b2) g g;  bc1 = 1 - b1^t;  bc2 = 1 - b2^t;
//     denom = sqrt(v) / sqrt(bc2) + eps;  p -= (lr / bc1) m / denom
// with the learning rate and the step count t read from device memory (schedulers and graph replays keep working).
#include "common.cuh"

namespace kdpc {

constexpr int ADAM_TENSORS = 96;         // per launch: 96 x 4 pointers + 97 chunk offsets = 3.5 KB of kernel parameters
constexpr int ADAM_CHUNK = 4096;         // elements per CTA
constexpr int ADAM_THREADS = 256;

struct AdamTable {
    float *p[ADAM_TENSORS];
    const float *g[ADAM_TENSORS];
    float *m[ADAM_TENSORS];
    float *v[ADAM_TENSORS];
    long long n[ADAM_TENSORS];
    int chunk0[ADAM_TENSORS + 1];        // first chunk of every tensor (prefix sums), chunk0[count] = number of chunks
    int count;
};

__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, float b1, float b2, float wd, float step_size,
                                         float bc2_sqrt, float eps) {
    g = wd != 0.f ? fmaf(wd, p, g) : g;
    m = fmaf(1.f - b1, g - m, m);                              // lerp
    v = fmaf(1.f - b2, g * g, v * b2);
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = fmaf(-step_size, m / denom, p);                        // addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(ADAM_THREADS)
adam_kernel(const __grid_constant__ AdamTable t, const float *__restrict__ lr, const float *__restrict__ step, float b1, float b2,
            float eps, float wd) {
    // which tensor does this chunk belong to (binary search over <= 97 prefix sums)
    int lo = 0, hi = t.count;
    const int c = blockIdx.x;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (t.chunk0[mid] <= c) lo = mid; else hi = mid;
    }
    const int ti = lo;
    const long long off = (long long)(c - t.chunk0[ti]) * ADAM_CHUNK;
    const long long n = t.n[ti];
    const int cnt = (int)min((long long)ADAM_CHUNK, n - off);
    float *p = t.p[ti] + off, *m = t.m[ti] + off, *v = t.v[ti] + off;
    const float *g = t.g[ti] + off;
    const float tt = __ldg(step) + 1.f;
    const float bc1 = 1.f - powf(b1, tt), bc2 = 1.f - powf(b2, tt);
    const float s = __ldg(lr) / bc1;                            // step_size (lr = 0: no update, nothing divides by it)
    const float bc2_sqrt_s = sqrtf(bc2), eps_s = eps;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec) {
        const int c4 = cnt >> 2;
        for (int i = threadIdx.x; i < c4; i += ADAM_THREADS) {
            float4 pp = reinterpret_cast<float4 *>(p)[i], mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
            const float4 gg = reinterpret_cast<const float4 *>(g)[i];
            adam_one(pp.x, gg.x, mm.x, vv.x, b1, b2, wd, s, bc2_sqrt_s, eps_s);
            adam_one(pp.y, gg.y, mm.y, vv.y, b1, b2, wd, s, bc2_sqrt_s, eps_s);
            adam_one(pp.z, gg.z, mm.z, vv.z, b1, b2, wd, s, bc2_sqrt_s, eps_s);
            adam_one(pp.w, gg.w, mm.w, vv.w, b1, b2, wd, s, bc2_sqrt_s, eps_s);
            reinterpret_cast<float4 *>(p)[i] = pp;
            reinterpret_cast<float4 *>(m)[i] = mm;
            reinterpret_cast<float4 *>(v)[i] = vv;
        }
        for (int i = (c4 << 2) + threadIdx.x; i < cnt; i += ADAM_THREADS) adam_one(p[i], g[i], m[i], v[i], b1, b2, wd, s, bc2_sqrt_s, eps_s);
    } else {
        for (int i = threadIdx.x; i < cnt; i += ADAM_THREADS) adam_one(p[i], g[i], m[i], v[i], b1, b2, wd, s, bc2_sqrt_s, eps_s);
    }
}

__global__ void adam_bump_kernel(float *step) { *step += 1.f; }

}  // namespace kdpc

using namespace kdpc;

/* One Adam step (torch.optim.Adam semantics: L2 weight decay, bias correction, no amsgrad) over `count` fp32 tensors.
 * params / grads / exp_avg / exp_avg_sq: HOST arrays of `count` device pointers, sizes: HOST array of element counts.
 * lr, step: DEVICE scalars (fp32); step is incremented after the update (t = step + 1 is used by it). */
KDPC_API int kdpc_adam_step(int count, const void *const *params, const void *const *grads, const void *const *exp_avg,
                            const void *const *exp_avg_sq, const long long *sizes, const float *lr, float beta1, float beta2,
                            float eps, float weight_decay, float *step, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(count >= 0 && (count == 0 || (params && grads && exp_avg && exp_avg_sq && sizes)) && lr && step);
    cudaStream_t st = to_stream(stream);
    for (int base = 0; base < count; base += ADAM_TENSORS) {
        AdamTable t;
        const int c = count - base < ADAM_TENSORS ? count - base : ADAM_TENSORS;
        long long chunks = 0;
        for (int i = 0; i < c; ++i) {
            KDPC_CHECK_ARGS(params[base + i] && grads[base + i] && exp_avg[base + i] && exp_avg_sq[base + i] && sizes[base + i] >= 0);
            t.p[i] = reinterpret_cast<float *>(const_cast<void *>(params[base + i]));
            t.g[i] = reinterpret_cast<const float *>(grads[base + i]);
            t.m[i] = reinterpret_cast<float *>(const_cast<void *>(exp_avg[base + i]));
            t.v[i] = reinterpret_cast<float *>(const_cast<void *>(exp_avg_sq[base + i]));
            t.n[i] = sizes[base + i];
            t.chunk0[i] = (int)chunks;
            chunks += (sizes[base + i] + ADAM_CHUNK - 1) / ADAM_CHUNK;
            if (chunks >= (1ll << 31)) return KDPC_EUNSUPPORTED;
        }
        for (int i = c; i <= ADAM_TENSORS; ++i) t.chunk0[i] = (int)chunks;
        t.count = c;
        if (chunks > 0) adam_kernel<<<(unsigned)chunks, ADAM_THREADS, 0, st>>>(t, lr, step, beta1, beta2, eps, weight_decay);
    }
    adam_bump_kernel<<<1, 1, 0, st>>>(step);
    KDPC_RETURN_LAST();
}
