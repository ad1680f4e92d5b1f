// Furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel<bs> (reference pointnet2/src/sampling_gpu.cu:93-253).
// Same arithmetic, same result bit-for-bit, different machine mapping:
//   * one CTA per cloud, the whole cloud staged ONCE in shared memory (SoA), every thread's
//     own points and their running min-distance held in registers for all M iterations
//     (the reference re-reads cloud + temp from global/L2 every iteration);
//   * arg-max by two REDUX.MAX per warp + one shared-memory hop + two REDUX.MAX, with ONE
//     __syncthreads per iteration (double-buffered slots) instead of a 10-level shared
//     memory tree with 11 barriers.
// Tie rule.  The reference's result depends on its block size bs = opt_n_threads(n)
// (cuda_utils.h:9-14): thread tid scans k = tid, tid+bs, ... keeping the FIRST strict
// maximum, and the left-biased tree (__update, sampling_gpu.cu:86-91) then prefers the
// smaller bit-reversed tid.  So the winner is the maximum of the key
//       ( d2 , - (bitrev(k mod bs) << 16 | k / bs) ).
// We launch exactly bs threads with the same k = tid + i*bs ownership, which makes the
// per-thread part of the rule free, and encode the rest in a 32+32 bit max-reduction.
#include "common.cuh"

namespace kdpc {

static inline int ref_block_size(int n) {   // cuda_utils.h:9-14
    int p = 1;
    while ((p << 1) <= n && (p << 1) <= 1024) p <<= 1;
    return p;
}

template <int PPT>
__global__ void __launch_bounds__(1024)
fps_smem_kernel(int n, int m, int lg, const float *__restrict__ xyz, float *__restrict__ temp,
                int *__restrict__ idx_out) {
    extern __shared__ float cloud[];          // x[n] | y[n] | z[n]
    __shared__ uint2 slots[2][32];
    float *sx = cloud, *sy = cloud + n, *sz = cloud + 2 * n;

    const int T = blockDim.x;                 // == 1 << lg  (>= 32)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const float *p = xyz + (size_t)blockIdx.x * n * 3;
    idx_out += (size_t)blockIdx.x * m;

    for (int i = tid; i < 3 * n; i += T) {    // coalesced AoS read -> SoA smem
        int k = i / 3, c = i - 3 * k;
        cloud[c * n + k] = p[i];
    }
    __syncthreads();

    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        int k = tid + i * T;
        bool v = k < n;
        px[i] = v ? sx[k] : 0.f;
        py[i] = v ? sy[k] : 0.f;
        pz[i] = v ? sz[k] : 0.f;
        md[i] = v ? 1e10f : -1.f;             // -1 never beats a real candidate (d2 >= 0)
    }
    const unsigned rtid = __brev((unsigned)tid) >> (32 - lg);

    int old = 0;
    if (tid == 0) idx_out[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = sx[old], y1 = sy[old], z1 = sz[old];
        float best = -1.f;
        int besti = 0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            float d = direct_dist(px[i] - x1, py[i] - y1, pz[i] - z1);
            float d2 = fminf(d, md[i]);
            md[i] = d2;
            bool gt = d2 > best;
            best = gt ? d2 : best;
            besti = gt ? i : besti;
        }
        const unsigned hi = __float_as_uint(best);
        const unsigned lo = ~((rtid << 16) | (unsigned)besti);
        const unsigned wh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned wl = __reduce_max_sync(0xffffffffu, hi == wh ? lo : 0u);
        if (lane == 0) slots[j & 1][warp] = make_uint2(wh, wl);
        __syncthreads();
        uint2 s = lane < nwarps ? slots[j & 1][lane] : make_uint2(0u, 0u);
        const unsigned gh = __reduce_max_sync(0xffffffffu, s.x);
        const unsigned gl = __reduce_max_sync(0xffffffffu, s.x == gh ? s.y : 0u);
        const unsigned key = ~gl;
        old = (int)(__brev(key >> 16) >> (32 - lg)) + (int)(key & 0xffffu) * T;
        if (tid == 0) idx_out[j] = old;
    }
    if (temp != nullptr) {
        temp += (size_t)blockIdx.x * n;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            int k = tid + i * T;
            if (k < n) temp[k] = md[i];
        }
    }
}

// Any n (n < 32, or a cloud too large for the register/shared-memory path): cloud and the
// min-distance field stay in global memory (L2 resident), 256 threads, same total order.
__global__ void __launch_bounds__(256)
fps_generic_kernel(int n, int m, int bs, int lg, const float *__restrict__ xyz, float *__restrict__ temp,
                   int *__restrict__ idx_out) {
    __shared__ uint2 slots[2][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *p = xyz + (size_t)blockIdx.x * n * 3;
    temp += (size_t)blockIdx.x * n;
    idx_out += (size_t)blockIdx.x * m;
    for (int k = tid; k < n; k += 256) temp[k] = 1e10f;
    __syncthreads();
    int old = 0;
    if (tid == 0) idx_out[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
        unsigned hi = 0u, lo = 0u;            // (0,0) = "no candidate"
        for (int k = tid; k < n; k += 256) {
            float d = direct_dist(p[k * 3 + 0] - x1, p[k * 3 + 1] - y1, p[k * 3 + 2] - z1);
            float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            unsigned r = lg ? (__brev((unsigned)k & (unsigned)(bs - 1)) >> (32 - lg)) : 0u;
            unsigned h = __float_as_uint(d2), l = ~((r << 16) | (unsigned)(k >> lg));
            bool better = h > hi || (h == hi && l > lo);
            hi = better ? h : hi;
            lo = better ? l : lo;
        }
        const unsigned wh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned wl = __reduce_max_sync(0xffffffffu, hi == wh ? lo : 0u);
        if (lane == 0) slots[j & 1][warp] = make_uint2(wh, wl);
        __syncthreads();
        uint2 s = lane < 8 ? slots[j & 1][lane] : make_uint2(0u, 0u);
        const unsigned gh = __reduce_max_sync(0xffffffffu, s.x);
        const unsigned gl = __reduce_max_sync(0xffffffffu, s.x == gh ? s.y : 0u);
        const unsigned key = ~gl;
        const unsigned r = key >> 16;
        old = (int)(lg ? (__brev(r) >> (32 - lg)) : 0u) + (int)(key & 0xffffu) * bs;
        if (tid == 0) idx_out[j] = old;
    }
}

template <int PPT>
static int launch_smem(int b, int n, int m, int bs, int lg, const float *xyz, float *temp, int *idx,
                       cudaStream_t st) {
    size_t smem = (size_t)3 * n * sizeof(float);
    KDPC_ENSURE_SMEM((fps_smem_kernel<PPT>), 3 * 8192 * (int)sizeof(float));     // largest cloud of the register path
    fps_smem_kernel<PPT><<<b, bs, smem, st>>>(n, m, lg, xyz, temp, idx);
    return (int)cudaGetLastError();
}

}  // namespace kdpc

KDPC_API int kdpc_fps(int b, int n, int m, const float *xyz, float *temp, int *idx, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(xyz && idx && b > 0 && n > 0);
    if (m <= 0) return KDPC_OK;                       // sampling_gpu.cu:100
    using namespace kdpc;
    cudaStream_t st = to_stream(stream);
    const int bs = ref_block_size(n);
    int lg = 0;
    while ((1 << lg) < bs) ++lg;
    const int ppt = (n + bs - 1) / bs;
    if (bs >= 32 && ppt <= 8) {
        if (ppt == 1) return launch_smem<1>(b, n, m, bs, lg, xyz, temp, idx, st);
        if (ppt == 2) return launch_smem<2>(b, n, m, bs, lg, xyz, temp, idx, st);
        if (ppt <= 4) return launch_smem<4>(b, n, m, bs, lg, xyz, temp, idx, st);
        return launch_smem<8>(b, n, m, bs, lg, xyz, temp, idx, st);
    }
    if (temp == nullptr || (n >> lg) >= 65536) return KDPC_EUNSUPPORTED;   // generic path needs the scratch field
    fps_generic_kernel<<<b, 256, 0, st>>>(n, m, bs, lg, xyz, temp, idx);
    KDPC_RETURN_LAST();
}
