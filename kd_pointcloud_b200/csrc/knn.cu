// Brute-force k-nearest-neighbour selection for sm_100a.
//
// Replaces  square_distance + torch.topk  (reference pointconv_util.py:73-107), which
// materialises a [B,S,N] fp32 matrix (2.1 GB at B=8, 8192x8192) and makes ~8 passes over it,
// and  three_nn_kernel_fast  (pointnet2/src/interpolate_gpu.cu:9-52).
//
// Mapping: one query per thread, its K best (distance, index) pairs in REGISTERS as a sorted
// list; candidates are pre-packed once per call to float4 (x, y, z, |c|^2) and streamed through
// shared memory in double-buffered tiles by 1-D bulk TMA (cp.async.bulk + mbarrier), every
// thread of the CTA reading the same candidate (shared-memory broadcast).  The [B,S,N] matrix
// never exists; HBM traffic is 12(S+N) + 16N + 4SK bytes per cloud.
//
// Selection cost is what matters (the kernel is issue-bound, not HBM-bound), so the hot loop
// only FILTERS against the thread's current K-th distance tau and appends survivors to a small
// per-thread queue in shared memory; the O(K) sorted insertion runs in a separate, rarely
// executed flush loop, which keeps warp divergence out of the distance loop.
//
// Order: ascending (distance, index) — candidates are visited in index order and only a
// strictly smaller distance displaces an entry.  torch.topk(sorted=False) leaves both the
// order and the choice among exact K-th ties unspecified; this fixes them (north_star).
// Distances use the exact rounding sequence of the reference expression (common.cuh).
#include "common.cuh"

namespace kdpc {

constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 512;        // candidates per smem stage (8 KB)
constexpr int KNN_CHUNK = 16;        // candidates between queue-pressure checks
constexpr int KNN_QCAP = 32;         // per-thread survivor queue capacity

constexpr int KNN_TILE_BYTES = 2 * KNN_TILE * 16;
constexpr int KNN_QUEUE_BYTES = KNN_QCAP * KNN_THREADS * 4;
constexpr int KNN_SMEM_BYTES = KNN_TILE_BYTES + 2 * KNN_QUEUE_BYTES;

enum { DIST_EXPANSION = 0, DIST_DIRECT = 1 };

__global__ void pack_cand4_kernel(long long total, const float *__restrict__ xyz, float4 *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float x = xyz[i * 3 + 0], y = xyz[i * 3 + 1], z = xyz[i * 3 + 2];
    out[i] = make_float4(x, y, z, sq_norm3(x, y, z));
}

template <int K>
__device__ __forceinline__ void sorted_insert(float (&ld)[K], int (&li)[K], float d, int i) {
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
        const bool shift = d < ld[j - 1];
        const bool here = d < ld[j];
        ld[j] = shift ? ld[j - 1] : (here ? d : ld[j]);
        li[j] = shift ? li[j - 1] : (here ? i : li[j]);
    }
    if (d < ld[0]) { ld[0] = d; li[0] = i; }
}

template <int K, int MODE>
__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(int s, int n, int k_out, const float *__restrict__ query, const float4 *__restrict__ cand4,
           int *__restrict__ idx32, long long *__restrict__ idx64, float *__restrict__ dist_out) {
    extern __shared__ __align__(128) unsigned char knn_smem[];
    float4 (*tile)[KNN_TILE] = reinterpret_cast<float4 (*)[KNN_TILE]>(knn_smem);
    float (*qd)[KNN_THREADS] = reinterpret_cast<float (*)[KNN_THREADS]>(knn_smem + KNN_TILE_BYTES);
    int (*qi)[KNN_THREADS] = reinterpret_cast<int (*)[KNN_THREADS]>(knn_smem + KNN_TILE_BYTES + KNN_QUEUE_BYTES);
    __shared__ __align__(8) uint64_t bar[2];

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int q = blockIdx.x * KNN_THREADS + tid;
    const bool active = q < s;
    const float4 *cb = cand4 + (size_t)b * n;
    const int ntiles = (n + KNN_TILE - 1) / KNN_TILE;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int t = 0; t < 2 && t < ntiles; ++t) {
            int cnt = min(KNN_TILE, n - t * KNN_TILE);
            mbar_expect_tx(&bar[t], cnt * 16);
            tma_load_1d(&tile[t][0], cb + (size_t)t * KNN_TILE, cnt * 16, &bar[t]);
        }
    }

    float qx = 0.f, qy = 0.f, qz = 0.f, qq = 0.f;
    if (active) {
        const float *qp = query + ((size_t)b * s + q) * 3;
        qx = qp[0]; qy = qp[1]; qz = qp[2];
        qq = sq_norm3(qx, qy, qz);
    }
    float ld[K];
    int li[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { ld[j] = INFINITY; li[j] = 0; }
    // inactive threads never pass the filter (tau = -inf)
    float tau = active ? INFINITY : -INFINITY;
    int qn = 0;

    auto flush = [&]() {
        const int mx = __reduce_max_sync(0xffffffffu, qn);
        for (int t = 0; t < mx; ++t) {
            if (t < qn) {
                const float d = qd[t][tid];
                if (d < tau) {
                    sorted_insert<K>(ld, li, d, qi[t][tid]);
                    tau = ld[K - 1];
                }
            }
        }
        qn = 0;
    };

    for (int t = 0; t < ntiles; ++t) {
        const int st = t & 1;
        const int cnt = min(KNN_TILE, n - t * KNN_TILE);
        const int base = t * KNN_TILE;
        mbar_wait(&bar[st], (t >> 1) & 1);
        const float4 *tp = tile[st];
        for (int j0 = 0; j0 < cnt; j0 += KNN_CHUNK) {
            const int jn = min(KNN_CHUNK, cnt - j0);
            if (jn == KNN_CHUNK) {
#pragma unroll
                for (int jj = 0; jj < KNN_CHUNK; ++jj) {
                    const float4 c = tp[j0 + jj];
                    const float d = MODE == DIST_EXPANSION
                                        ? expansion_dist(qx, qy, qz, qq, c.x, c.y, c.z, c.w)
                                        : direct_dist(qx - c.x, qy - c.y, qz - c.z);
                    if (d < tau) { qd[qn][tid] = d; qi[qn][tid] = base + j0 + jj; ++qn; }
                }
            } else {
                for (int jj = 0; jj < jn; ++jj) {
                    const float4 c = tp[j0 + jj];
                    const float d = MODE == DIST_EXPANSION
                                        ? expansion_dist(qx, qy, qz, qq, c.x, c.y, c.z, c.w)
                                        : direct_dist(qx - c.x, qy - c.y, qz - c.z);
                    if (d < tau) { qd[qn][tid] = d; qi[qn][tid] = base + j0 + jj; ++qn; }
                }
            }
            if (__any_sync(0xffffffffu, qn > KNN_QCAP - KNN_CHUNK)) flush();
        }
        __syncthreads();                       // everyone is done reading tile[st]
        if (tid == 0 && t + 2 < ntiles) {
            int c2 = min(KNN_TILE, n - (t + 2) * KNN_TILE);
            mbar_expect_tx(&bar[st], c2 * 16);
            tma_load_1d(&tile[st][0], cb + (size_t)(t + 2) * KNN_TILE, c2 * 16, &bar[st]);
        }
    }
    flush();

    if (active) {
        const size_t o = ((size_t)b * s + q) * k_out;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (j < k_out) {
                if (idx32) idx32[o + j] = li[j];
                if (idx64) idx64[o + j] = li[j];
                if (dist_out) dist_out[o + j] = ld[j];
            }
        }
    }
}

template <int MODE>
static int launch_knn(int b, int s, int n, int k, const float *query, const float *cand, void *ws,
                      int *idx32, long long *idx64, float *dist, cudaStream_t st) {
    float4 *c4 = reinterpret_cast<float4 *>(ws);
    const long long total = (long long)b * n;
    pack_cand4_kernel<<<(unsigned)div_up_ll(total, 256), 256, 0, st>>>(total, cand, c4);
    dim3 grid((s + KNN_THREADS - 1) / KNN_THREADS, b);
#define KDPC_KNN_CASE(KT) \
    if (k <= KT) { \
        KDPC_ENSURE_SMEM((knn_kernel<KT, MODE>), KNN_SMEM_BYTES); \
        knn_kernel<KT, MODE><<<grid, KNN_THREADS, KNN_SMEM_BYTES, st>>>(s, n, k, query, c4, idx32, idx64, dist); \
        return (int)cudaGetLastError(); }
    KDPC_KNN_CASE(1)
    KDPC_KNN_CASE(3)
    KDPC_KNN_CASE(5)
    KDPC_KNN_CASE(9)
    KDPC_KNN_CASE(10)
    KDPC_KNN_CASE(16)
    KDPC_KNN_CASE(24)
    KDPC_KNN_CASE(32)
#undef KDPC_KNN_CASE
    return KDPC_EUNSUPPORTED;
}

__global__ void square_distance_kernel(int s, int n, const float *__restrict__ src, const float *__restrict__ dst,
                                       float *__restrict__ out) {
    const int b = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float *qp = src + ((size_t)b * s + i) * 3;
    const float *cp = dst + ((size_t)b * n + j) * 3;
    float qx = qp[0], qy = qp[1], qz = qp[2], cx = cp[0], cy = cp[1], cz = cp[2];
    out[((size_t)b * s + i) * n + j] =
        expansion_dist(qx, qy, qz, sq_norm3(qx, qy, qz), cx, cy, cz, sq_norm3(cx, cy, cz));
}

}  // namespace kdpc

// Workspace of kdpc_knn / kdpc_three_nn: two spatially sorted clouds (knn_bf.cu) when the pruned search
// applies, the float4-packed candidates of the brute-force kernel otherwise.
static inline bool use_pruned(int s, int n) { return n >= 256 && n <= 16384 && s <= 16384; }

KDPC_API long long kdpc_knn_workspace_bytes(int b, int s, int n) {
    if (b <= 0 || s <= 0 || n <= 0) return 0;
    if (use_pruned(s, n)) return kdpc_spatial_sort_bytes(b, n) + kdpc_spatial_sort_bytes(b, s);
    return (long long)b * n * 16;
}

static int knn_dispatch(int b, int s, int n, int k, int direct, const float *query, const float *cand, void *ws,
                        int *idx32, long long *idx64, float *dist, kdpc_stream_t stream) {
    if (k > 32 || b > 65535) return KDPC_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(ws) % 16) != 0) return KDPC_EINVAL;
    if (use_pruned(s, n) && k <= n) {
        void *cws = ws;
        void *qws = reinterpret_cast<unsigned char *>(ws) + kdpc_spatial_sort_bytes(b, n);
        int rc = kdpc_spatial_sort(b, n, cand, cws, stream);
        if (rc != 0) return rc;
        if (query == cand && s == n) {
            qws = cws;
        } else {
            rc = kdpc_spatial_sort(b, s, query, qws, stream);
            if (rc != 0) return rc;
        }
        return kdpc_knn_sorted(b, s, n, k, direct, qws, cws, idx32, idx64, dist, stream);
    }
    if (direct)
        return kdpc::launch_knn<kdpc::DIST_DIRECT>(b, s, n, k, query, cand, ws, idx32, idx64, dist, to_stream(stream));
    return kdpc::launch_knn<kdpc::DIST_EXPANSION>(b, s, n, k, query, cand, ws, idx32, idx64, dist, to_stream(stream));
}

KDPC_API int kdpc_knn(int b, int s, int n, int k, const float *query, const float *cand, void *ws,
                      int *idx32, long long *idx64, float *dist, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(query && cand && ws && b > 0 && s > 0 && n > 0 && k > 0);
    return knn_dispatch(b, s, n, k, 0, query, cand, ws, idx32, idx64, dist, stream);
}

KDPC_API int kdpc_knn_bruteforce(int b, int s, int n, int k, const float *query, const float *cand, void *ws,
                                 int *idx32, long long *idx64, float *dist, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(query && cand && ws && b > 0 && s > 0 && n > 0 && k > 0);
    if (k > 32 || b > 65535) return KDPC_EUNSUPPORTED;
    return kdpc::launch_knn<kdpc::DIST_EXPANSION>(b, s, n, k, query, cand, ws, idx32, idx64, dist, to_stream(stream));
}

KDPC_API int kdpc_three_nn(int b, int n, int m, const float *unknown, const float *known, void *ws,
                           float *dist2, int *idx, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(unknown && known && ws && dist2 && idx && b > 0 && n > 0 && m > 0);
    return knn_dispatch(b, n, m, 3, 1, unknown, known, ws, idx, nullptr, dist2, stream);
}

KDPC_API int kdpc_square_distance(int b, int s, int n, const float *src, const float *dst, float *out,
                                  kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(src && dst && out && b > 0 && s > 0 && n > 0);
    if (b > 65535 || s > 65535) return KDPC_EUNSUPPORTED;
    dim3 grid((n + 255) / 256, s, b);
    kdpc::square_distance_kernel<<<grid, 256, 0, to_stream(stream)>>>(s, n, src, dst, out);
    KDPC_RETURN_LAST();
}
