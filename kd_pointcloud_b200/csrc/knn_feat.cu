// k nearest neighbours in a C-dimensional FEATURE space (C = 4 .. 512), sm_100a.
//
// Replaces  square_distance + torch.topk  on feature tensors: CrossLayerLightFG picks half of its cost-volume
// neighbourhood by feature distance (reference pointconv_util.py:1905  knn_point(nsample//2, knn2, knn1) with
// knn1/knn2 = [B,N,C] features, via square_distance :73-94 and topk :106).  The reference materialises the [B,S,N]
// matrix with an sgemm; here one thread owns one query and its K best in registers, candidates stream through shared
// memory in tiles of 32 (every thread reads the same candidate value: a broadcast), the query row is walked in chunks
// of 32 dims with 32 running dot products in registers.  CUDA cores only: this is the variant models' path
// (models_bid_FG.py, models_bifeat.py), not the headline's.
//
// Distance = |q|^2 + |c|^2 - 2 q.c evaluated as the reference's expression rn(rn(-2*dot + |q|^2) + |c|^2); the
// dot / norm summation order is ascending in the channel index (cuBLAS' order is unspecified, so K-th-boundary
// agreement with the reference is at rounding level, as for the coordinate kNN).  Order: ascending (distance, index).
#include "common.cuh"

namespace kdpc {

constexpr int KFT_THREADS = 128;     // queries per CTA
constexpr int KFT_TILE = 32;         // candidates per shared-memory tile
constexpr int KFT_DCH = 32;          // channels per register chunk

template <int K>
__device__ __forceinline__ void kft_insert(float (&ld)[K], int (&li)[K], float d, int i) {
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
        const bool shift = d < ld[j - 1];
        const bool here = d < ld[j];
        ld[j] = shift ? ld[j - 1] : (here ? d : ld[j]);
        li[j] = shift ? li[j - 1] : (here ? i : li[j]);
    }
    if (d < ld[0]) { ld[0] = d; li[0] = i; }
}

template <int K>
__global__ void __launch_bounds__(KFT_THREADS)
knn_feat_kernel(int s, int n, int c, int cpad, int k_out, const float *__restrict__ query, const float *__restrict__ cand,
                int *__restrict__ idx, float *__restrict__ dist_out) {
    extern __shared__ __align__(16) float kft_smem[];            // tile [KFT_TILE][cpad] | norms [KFT_TILE]
    float *tile = kft_smem;
    float *cnorm = kft_smem + (size_t)KFT_TILE * cpad;
    const int b = blockIdx.y;
    const int q = blockIdx.x * KFT_THREADS + threadIdx.x;
    const bool active = q < s;
    const float *qrow = query + ((size_t)b * s + (active ? q : 0)) * c;
    const float *cb = cand + (size_t)b * n * c;

    float qq = 0.f;
    for (int d = 0; d < c; ++d) { const float v = __ldg(qrow + d); qq = __fmaf_rn(v, v, qq); }

    float ld[K];
    int li[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { ld[j] = 3.0e38f; li[j] = 0; }

    for (int t0 = 0; t0 < n; t0 += KFT_TILE) {
        const int tn = min(KFT_TILE, n - t0);
        __syncthreads();
        // stage the candidate tile (rows padded with zeros to cpad) and its squared norms
        for (int e = threadIdx.x; e < KFT_TILE * cpad; e += KFT_THREADS) {
            const int j = e / cpad, d = e - j * cpad;
            tile[e] = (j < tn && d < c) ? __ldg(cb + (size_t)(t0 + j) * c + d) : 0.f;
        }
        __syncthreads();
        if (threadIdx.x < KFT_TILE) {
            float cc = 0.f;
            const float *r = tile + threadIdx.x * cpad;
            for (int d = 0; d < c; ++d) cc = __fmaf_rn(r[d], r[d], cc);
            cnorm[threadIdx.x] = cc;
        }
        __syncthreads();
        float acc[KFT_TILE];
#pragma unroll
        for (int j = 0; j < KFT_TILE; ++j) acc[j] = 0.f;
        for (int d0 = 0; d0 < cpad; d0 += KFT_DCH) {
            float qv[KFT_DCH];
#pragma unroll
            for (int d = 0; d < KFT_DCH; ++d) qv[d] = (d0 + d < c) ? __ldg(qrow + d0 + d) : 0.f;
#pragma unroll
            for (int j = 0; j < KFT_TILE; ++j) {
                const float4 *r4 = reinterpret_cast<const float4 *>(tile + j * cpad + d0);
#pragma unroll
                for (int d4 = 0; d4 < KFT_DCH / 4; ++d4) {
                    const float4 cv = r4[d4];                    // same address for the whole warp: broadcast
                    acc[j] = __fmaf_rn(qv[4 * d4 + 0], cv.x, acc[j]);
                    acc[j] = __fmaf_rn(qv[4 * d4 + 1], cv.y, acc[j]);
                    acc[j] = __fmaf_rn(qv[4 * d4 + 2], cv.z, acc[j]);
                    acc[j] = __fmaf_rn(qv[4 * d4 + 3], cv.w, acc[j]);
                }
            }
        }
        if (active) {
#pragma unroll
            for (int j = 0; j < KFT_TILE; ++j) {
                if (j < tn) {
                    const float dd = __fadd_rn(__fmaf_rn(-2.f, acc[j], qq), cnorm[j]);
                    if (dd < ld[K - 1]) kft_insert<K>(ld, li, dd, t0 + j);
                }
            }
        }
    }
    if (active) {
        int *o = idx + ((size_t)b * s + q) * k_out;
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (j < k_out) o[j] = li[j];
        if (dist_out != nullptr) {
            float *od = dist_out + ((size_t)b * s + q) * k_out;
#pragma unroll
            for (int j = 0; j < K; ++j)
                if (j < k_out) od[j] = ld[j];
        }
    }
}

}  // namespace kdpc

using namespace kdpc;

/* query [B,S,C], cand [B,N,C] -> idx int32 [B,S,k] (ascending (distance, index)), dist [B,S,k] or NULL.  1 <= k <= 32 <= N. */
KDPC_API int kdpc_knn_feat(int b, int s, int n, int c, int k, const float *query, const float *cand, int *idx, float *dist,
                           kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(query && cand && idx && b > 0 && s > 0 && n > 0 && c > 0 && k > 0);
    if (k > 32 || k > n || c > 512 || b > 65535) return KDPC_EUNSUPPORTED;
    const int cpad = (c + KFT_DCH - 1) / KFT_DCH * KFT_DCH;
    const size_t smem = ((size_t)KFT_TILE * cpad + KFT_TILE) * sizeof(float);
    dim3 grid((s + KFT_THREADS - 1) / KFT_THREADS, b);
    cudaStream_t st = to_stream(stream);
    if (k <= 8) {
        KDPC_ENSURE_SMEM(knn_feat_kernel<8>, 80 * 1024);
        knn_feat_kernel<8><<<grid, KFT_THREADS, smem, st>>>(s, n, c, cpad, k, query, cand, idx, dist);
    } else if (k <= 16) {
        KDPC_ENSURE_SMEM(knn_feat_kernel<16>, 80 * 1024);
        knn_feat_kernel<16><<<grid, KFT_THREADS, smem, st>>>(s, n, c, cpad, k, query, cand, idx, dist);
    } else {
        KDPC_ENSURE_SMEM(knn_feat_kernel<32>, 80 * 1024);
        knn_feat_kernel<32><<<grid, KFT_THREADS, smem, st>>>(s, n, c, cpad, k, query, cand, idx, dist);
    }
    KDPC_RETURN_LAST();
}
