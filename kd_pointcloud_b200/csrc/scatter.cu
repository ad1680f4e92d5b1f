// Deterministic scatter-add for the backward of every gather-type op.
//
// The reference accumulates gradients with fp32 atomicAdd (pointnet2/src/sampling_gpu.cu:62,
// group_points_gpu.cu:24, interpolate_gpu.cu:139-141): order-dependent, hence run-to-run
// different sums, and heavily contended (each point is selected ~K times).  Here the index
// list is inverted ONCE into a CSR ("who selected me", members ascending), and every gradient
// is then a segmented gather-reduce in a fixed order: no atomics on floats, coalesced, and
// the CSR is shared by all layers that reuse the same index tensor.
#include "common.cuh"

namespace kdpc {

constexpr int CSR_THREADS = 1024;
constexpr int CSR_LONG = 512;                             // longest segment sorted through shared memory

// CTA (part, b) inverts the candidates i in [i_lo, i_hi) of cloud b: it scans the WHOLE index list of the cloud (1 MB at
// most, L2-resident) but counts / places only the entries that select its candidates, plus one scalar - how many
// entries select a smaller candidate - which is its segments' base offset.  No communication between the parts, so a
// cloud is spread over `parts` SMs; the per-candidate segments are then sorted ascending (deterministic order) with
// one thread per segment.  (One CTA per cloud spent ~65 of its ~127 us insertion-sorting 8 segments per thread in global
// memory while 130 SMs idled: 26 builds per KD step.)
__global__ void __launch_bounds__(CSR_THREADS)
build_csr_kernel(int n, int m, int per_part, const int *__restrict__ idx, int *__restrict__ offsets, int *__restrict__ perm) {
    extern __shared__ int cnt[];                          // [per_part] counters, then CSR_LONG ints of scratch per warp
    int *scratch = cnt + ((per_part + 3) & ~3);
    __shared__ int warp_sums[32];
    __shared__ int below_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i_lo = blockIdx.x * per_part, i_hi = min(n, i_lo + per_part), span = i_hi - i_lo;
    idx += (size_t)blockIdx.y * m;
    offsets += (size_t)blockIdx.y * (n + 1);
    perm += (size_t)blockIdx.y * m;

    for (int i = tid; i < span; i += CSR_THREADS) cnt[i] = 0;
    if (tid == 0) below_s = 0;
    __syncthreads();
    int below = 0;
    for (int j = tid; j < m; j += CSR_THREADS) {
        const int v = idx[j];
        if (v < i_lo) ++below;
        else if (v < i_hi) atomicAdd(&cnt[v - i_lo], 1);  // integer: order-free
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if (lane == 0 && below) atomicAdd(&below_s, below);
    __syncthreads();

    // exclusive scan: contiguous chunk per thread, then warp + block scan of the chunk sums
    const int chunk = (span + CSR_THREADS - 1) / CSR_THREADS;
    const int lo = min(span, tid * chunk), hi = min(span, lo + chunk);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += cnt[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        warp_sums[lane] = wi - w;                         // exclusive prefix of warp totals
    }
    __syncthreads();
    int run = below_s + warp_sums[warp] + incl - sum;
    for (int i = lo; i < hi; ++i) {
        const int c = cnt[i];
        offsets[i_lo + i] = run;
        cnt[i] = run;                                     // becomes the fill cursor
        run += c;
    }
    if (i_hi == n && tid == CSR_THREADS - 1) offsets[n] = m;
    __syncthreads();

    for (int j = tid; j < m; j += CSR_THREADS) {
        const int v = idx[j];
        if (v >= i_lo && v < i_hi) perm[atomicAdd(&cnt[v - i_lo], 1)] = j;
    }
    __syncthreads();
    // cursors now hold segment ends; sort each segment ascending => deterministic order.  One WARP per segment: up to 128
    // members in registers (four per lane), bitonic network over shuffles (strides < 32) and in-lane exchanges (strides 32,
    // 64).  (One thread per segment insertion-sorting in global memory left 95 % of the CTA idle and took up to 410 us for
    // the 3-NN index lists, whose segments are few and uneven: 1.8 ms of a 30 ms KD step over 26 builds.)
    for (int i = warp; i < span; i += CSR_THREADS / 32) {
        const int e = cnt[i];
        const int s0 = offsets[i_lo + i];
        const int len = e - s0;
        if (len <= 1) continue;
        if (len <= 128) {
            int v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = (q * 32 + lane < len) ? perm[s0 + q * 32 + lane] : 0x7fffffff;
            const int top = len <= 32 ? 32 : (len <= 64 ? 64 : 128);       // network size (warp-uniform)
            for (int k = 2; k <= top; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    int nv[4];                                              // (all four from the OLD values: in-lane partners)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int el = q * 32 + lane;                       // element index of v[q]
                        const int other = j < 32 ? __shfl_xor_sync(0xffffffffu, v[q], j) : v[q ^ (j >> 5)];
                        const bool up = (el & k) == 0;                      // ascending block
                        const bool lower = (el & j) == 0;                   // this element is the lower partner
                        nv[q] = (up == lower) ? min(v[q], other) : max(v[q], other);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = nv[q];
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q * 32 + lane < len) perm[s0 + q * 32 + lane] = v[q];
        } else if (len <= CSR_LONG) {
            // long segments (3-NN lists: a sparse point at the rim of a cloud is the neighbour of hundreds of dense points -
            // 453 members measured): rank sort through this warp's shared-memory scratch, len^2 / 32 compares per lane
            int *sc = scratch + warp * CSR_LONG;
            for (int t = lane; t < len; t += 32) sc[t] = perm[s0 + t];
            __syncwarp();
            for (int t = lane; t < len; t += 32) {
                const int x = sc[t];
                int rank = 0;
                for (int u = 0; u < len; ++u) rank += sc[u] < x ? 1 : 0;    // (members are distinct entry indices)
                perm[s0 + rank] = x;
            }
            __syncwarp();
        } else if (lane == 0) {                                            // (longer still: insertion sort, one lane)
            for (int a = s0 + 1; a < e; ++a) {
                const int v = perm[a];
                int p = a - 1;
                while (p >= s0 && perm[p] > v) { perm[p + 1] = perm[p]; --p; }
                perm[p + 1] = v;
            }
        }
    }
}

// point-major: grad_f[b,i,:] = sum_{p in seg(i)} wgt[b,perm[p]] * g[b, perm[p] / gdiv, :]
template <int VEC, bool ACCUM>
__global__ void __launch_bounds__(256)
scatter_rows_csr_kernel(long long total, int n, int m, int cvec, int gdiv, const float *__restrict__ g,
                        const float *__restrict__ wgt, const int *__restrict__ offsets, const int *__restrict__ perm,
                        float *__restrict__ grad_f) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long r = e / cvec;                         // b*n + i
    const int cv = (int)(e - r * cvec);
    const long long b = r / n;
    const int i = (int)(r - b * n);
    const int *off = offsets + b * (n + 1);
    const int *pm = perm + b * m;
    const float *wb = wgt ? wgt + b * m : nullptr;
    const int s0 = off[i], s1 = off[i + 1];
    const size_t grows = (size_t)(m / gdiv);
    if (VEC == 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 *g4 = reinterpret_cast<const float4 *>(g) + (size_t)b * grows * cvec;
        for (int p = s0; p < s1; ++p) {
            const int j = pm[p];
            const float w = wb ? wb[j] : 1.f;
            const float4 v = __ldg(g4 + (size_t)(j / gdiv) * cvec + cv);
            acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
        }
        float4 *o = reinterpret_cast<float4 *>(grad_f) + e;
        if (ACCUM) { float4 c = *o; acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w; }
        *o = acc;
    } else {
        float acc = 0.f;
        const float *gb = g + (size_t)b * grows * cvec;
        for (int p = s0; p < s1; ++p) {
            const int j = pm[p];
            const float w = wb ? wb[j] : 1.f;
            acc += w * __ldg(gb + (size_t)(j / gdiv) * cvec + cv);
        }
        if (ACCUM) acc += grad_f[e];
        grad_f[e] = acc;
    }
}

// channel-major (pointnet2 API): grad_f[b,c,i] = sum_{p in seg(i)} wgt[b,perm[p]] * g[b,c,perm[p]/gdiv]
__global__ void __launch_bounds__(256)
scatter_cm_csr_kernel(int c, int n, int m, int gdiv, const float *__restrict__ g, const float *__restrict__ wgt,
                      const int *__restrict__ offsets, const int *__restrict__ perm, float *__restrict__ grad_f) {
    const int b = blockIdx.z, ci = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int *off = offsets + (size_t)b * (n + 1);
    const int *pm = perm + (size_t)b * m;
    const float *wb = wgt ? wgt + (size_t)b * m : nullptr;
    const int gcols = m / gdiv;
    const float *gr = g + ((size_t)b * c + ci) * gcols;
    float acc = 0.f;
    for (int p = off[i]; p < off[i + 1]; ++p) {
        const int j = pm[p];
        acc += (wb ? wb[j] : 1.f) * __ldg(gr + j / gdiv);
    }
    grad_f[((size_t)b * c + ci) * n + i] = acc;
}

static int build_csr(int b, int n, int m, const int *idx, int *offsets, int *perm, cudaStream_t st) {
    if (b > 65535) return KDPC_EUNSUPPORTED;
    // ~2 CTAs per SM over all clouds, at least 64 candidates per part (and one segment per thread when possible)
    int parts = (2 * device_sms() + b - 1) / b;
    if (parts > (n + 63) / 64) parts = (n + 63) / 64;
    if (parts < 1) parts = 1;
    const int per_part = (n + parts - 1) / parts;
    parts = (n + per_part - 1) / per_part;
    const size_t smem = ((size_t)((per_part + 3) & ~3) + (size_t)(CSR_THREADS / 32) * CSR_LONG) * sizeof(int);
    if (smem > 200 * 1024) return KDPC_EUNSUPPORTED;
    KDPC_ENSURE_SMEM(build_csr_kernel, 200 * 1024);
    dim3 grid(parts, b);
    build_csr_kernel<<<grid, CSR_THREADS, smem, st>>>(n, m, per_part, idx, offsets, perm);
    return (int)cudaGetLastError();
}

static int scatter_cm(int b, int c, int n, int m, int gdiv, const float *g, const float *wgt, const int *idx,
                      void *ws, float *grad_f, cudaStream_t st) {
    if (b > 65535 || c > 65535) return KDPC_EUNSUPPORTED;
    int *offsets = reinterpret_cast<int *>(ws);
    int *perm = offsets + (size_t)b * (n + 1);
    int rc = build_csr(b, n, m, idx, offsets, perm, st);
    if (rc != 0) return rc;
    dim3 grid((n + 255) / 256, c, b);
    scatter_cm_csr_kernel<<<grid, 256, 0, st>>>(c, n, m, gdiv, g, wgt, offsets, perm, grad_f);
    return (int)cudaGetLastError();
}

}  // namespace kdpc

using namespace kdpc;

KDPC_API int kdpc_build_csr(int b, int n, int m, const int *idx, int *offsets, int *perm, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(idx && offsets && perm && b > 0 && n > 0 && m > 0);
    return build_csr(b, n, m, idx, offsets, perm, to_stream(stream));
}

KDPC_API int kdpc_scatter_rows_csr(int b, int n, int m, int c, int gdiv, const float *g, const float *wgt,
                                   const int *offsets, const int *perm, float *grad_f, int accumulate,
                                   kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(g && offsets && perm && grad_f && b > 0 && n > 0 && m > 0 && c > 0 && gdiv > 0);
    cudaStream_t st = to_stream(stream);
    const bool vec = (c % 4 == 0) && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(grad_f)) % 16 == 0);
    const int cvec = vec ? c / 4 : c;
    const long long total = (long long)b * n * cvec;
    const unsigned grid = (unsigned)div_up_ll(total, 256);
    if (vec) {
        if (accumulate) scatter_rows_csr_kernel<4, true><<<grid, 256, 0, st>>>(total, n, m, cvec, gdiv, g, wgt, offsets, perm, grad_f);
        else scatter_rows_csr_kernel<4, false><<<grid, 256, 0, st>>>(total, n, m, cvec, gdiv, g, wgt, offsets, perm, grad_f);
    } else {
        if (accumulate) scatter_rows_csr_kernel<1, true><<<grid, 256, 0, st>>>(total, n, m, cvec, gdiv, g, wgt, offsets, perm, grad_f);
        else scatter_rows_csr_kernel<1, false><<<grid, 256, 0, st>>>(total, n, m, cvec, gdiv, g, wgt, offsets, perm, grad_f);
    }
    KDPC_RETURN_LAST();
}

KDPC_API int kdpc_gather_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, void *ws,
                              float *grad_f, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(grad_out && idx && ws && grad_f && b > 0 && c > 0 && n > 0 && m > 0);
    return scatter_cm(b, c, n, m, 1, grad_out, nullptr, idx, ws, grad_f, to_stream(stream));
}

KDPC_API int kdpc_group_grad(int b, int c, int n, int s, int k, const float *grad_out, const int *idx, void *ws,
                             float *grad_f, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(grad_out && idx && ws && grad_f && b > 0 && c > 0 && n > 0 && s > 0 && k > 0);
    if ((long long)s * k > 0x7fffffffLL) return KDPC_EUNSUPPORTED;
    return scatter_cm(b, c, n, s * k, 1, grad_out, nullptr, idx, ws, grad_f, to_stream(stream));
}

KDPC_API int kdpc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                         const float *w, void *ws, float *grad_f, kdpc_stream_t stream) {
    // reference argument order (interpolate_gpu.h): n = number of interpolated points, m = source points
    KDPC_CHECK_ARGS(grad_out && idx && w && ws && grad_f && b > 0 && c > 0 && n > 0 && m > 0);
    return scatter_cm(b, c, m, n * 3, 3, grad_out, w, idx, ws, grad_f, to_stream(stream));
}
