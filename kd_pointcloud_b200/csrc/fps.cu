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

// ---- cluster version: ONE cloud spread over a cluster of 8 CTAs (8 SMs) ----------------------------------
// The single-CTA kernel is ISSUE-bound (32 warps x ~110 instructions per iteration on one SM, ~1500 cycles)
// while 132+ SMs idle.  Here 8 CTAs x 256 threads share a cloud: every thread keeps n/2048 points in
// registers, every CTA keeps the whole cloud in its shared memory (to read the winner's coordinates), and the
// per-iteration arg-max is ONE exchange: each warp's (value, tie-key) goes straight to all 8 CTAs through
// distributed shared memory (one 8-byte remote store per peer, double buffered, a parity tag in the unused sign
// bit of the distance; readers poll their LOCAL shared memory), then every CTA reduces the 64 warp results
// redundantly.  No barrier of any kind inside the loop.
// Ownership keeps the reference's tie rule exact: the 2048 threads are (part, reference thread t = gtid % bs);
// a thread owns k = t + i*bs for a contiguous range of i, scanned in ascending i with a strict '>'.
constexpr int FPSC_TOTAL = 2048;                               // default threads per cloud, over 8 x 256 or 4 x 512

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_remote_v2(uint32_t remote_addr, unsigned a, unsigned b) {     // one 8-byte transaction
    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(remote_addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_v2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.volatile.shared::cta.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}

template <int PPT, int FPSC_CTAS, int TOTAL>
__global__ void __launch_bounds__(TOTAL / FPSC_CTAS)
fps_cluster_kernel(int n, int m, int bs, int lg, const float *__restrict__ xyz, float *__restrict__ temp,
                   int *__restrict__ idx_out) {
    constexpr int NWARPS = TOTAL / 32;                         // warp results per iteration: 64 (two per lane) or 32
    extern __shared__ float cloud[];          // x[n] | y[n] | z[n]
    __shared__ __align__(8) uint2 slots[2][NWARPS];
    float *sx = cloud, *sy = cloud + n, *sz = cloud + 2 * n;
    constexpr int FPSC_THREADS = TOTAL / FPSC_CTAS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = cluster_ctarank();
    const int cloud_id = blockIdx.x / FPSC_CTAS;
    const int gtid = (int)rank * FPSC_THREADS + tid;           // 0 .. TOTAL-1
    const float *p = xyz + (size_t)cloud_id * n * 3;
    idx_out += (size_t)cloud_id * m;

    for (int i = tid; i < 2 * NWARPS; i += FPSC_THREADS) (&slots[0][0])[i] = make_uint2(0u, 0u);      // tag 0 = "nothing yet"
    for (int i = tid; i < 3 * n; i += FPSC_THREADS) {
        int k = i / 3, c = i - 3 * k;
        cloud[c * n + k] = p[i];
    }
    __syncthreads();
    cluster_sync_all();                                        // every CTA's slots are cleared before anyone sends

    const int t = gtid & (bs - 1);                             // reference thread
    const int i0 = (gtid >> lg) * PPT;                         // first owned i
    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = t + (i0 + i) * bs;
        const bool v = k < n;
        px[i] = v ? sx[k] : 0.f;
        py[i] = v ? sy[k] : 0.f;
        pz[i] = v ? sz[k] : 0.f;
        md[i] = v ? 1e10f : -1.f;
    }
    const unsigned rt = __brev((unsigned)t) >> (32 - lg);
    // lanes 0..CTAS-1 forward this warp's result to CTA `lane`: remote slot addresses (per buffer)
    uint32_t rslot[2] = {0, 0};
    if (lane < FPSC_CTAS) {
#pragma unroll
        for (int par = 0; par < 2; ++par)
            rslot[par] = map_to_cta(smem_u32(&slots[par][rank * (FPSC_THREADS / 32) + warp]), (unsigned)lane);
    }

    int old = 0;
    if (gtid == 0) idx_out[0] = 0;
    for (int j = 1; j < m; ++j) {
        const int par = j & 1;
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
        const unsigned lo = ~((rt << 16) | (unsigned)(i0 + besti));
        const unsigned wh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned wl = __reduce_max_sync(0xffffffffu, hi == wh ? lo : 0u);
        // bit 31 of the value word (distances are >= 0) carries a use-parity tag, so that readers can tell this
        // iteration's result from the one written two iterations ago into the same (double-buffered) slot
        const unsigned tag = ((((unsigned)(j - 1) >> 1) + 1u) & 1u) << 31;
        if (lane < FPSC_CTAS) st_remote_v2(rslot[par], wh | tag, wl);
        unsigned mh, ml;
        if constexpr (NWARPS == 64) {
            uint2 s0, s1;
            const uint32_t a0 = smem_u32(&slots[par][lane]), a1 = smem_u32(&slots[par][lane + 32]);
            unsigned spins = 0;
            for (;;) {                                         // poll local shared memory until both slots are fresh
                s0 = ld_volatile_v2(a0);
                s1 = ld_volatile_v2(a1);
                if (((s0.x ^ tag) | (s1.x ^ tag)) >> 31 == 0u) break;
                if (++spins > (1u << 22)) __trap();            // a lost peer: fail the launch instead of hanging
            }
            s0.x &= 0x7fffffffu;
            s1.x &= 0x7fffffffu;
            const bool second = s1.x > s0.x || (s1.x == s0.x && s1.y > s0.y);
            mh = second ? s1.x : s0.x;
            ml = second ? s1.y : s0.y;
        } else {
            static_assert(NWARPS == 32 || NWARPS == 64, "one or two warp results per lane");
            uint2 s0;
            const uint32_t a0 = smem_u32(&slots[par][lane]);
            unsigned spins = 0;
            for (;;) {
                s0 = ld_volatile_v2(a0);
                if ((s0.x ^ tag) >> 31 == 0u) break;
                if (++spins > (1u << 22)) __trap();
            }
            mh = s0.x & 0x7fffffffu;
            ml = s0.y;
        }
        const unsigned gh = __reduce_max_sync(0xffffffffu, mh);
        const unsigned gl = __reduce_max_sync(0xffffffffu, mh == gh ? ml : 0u);
        const unsigned key = ~gl;
        old = (int)(__brev(key >> 16) >> (32 - lg)) + (int)(key & 0xffffu) * bs;
        if (gtid == 0) idx_out[j] = old;
    }
    if (temp != nullptr) {
        temp += (size_t)cloud_id * n;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = t + (i0 + i) * bs;
            if (k < n) temp[k] = md[i];
        }
    }
    cluster_sync_all();                                        // nobody exits while a peer may still write to it
}

static int fps_cluster_spread = 0;                             // A/B: cudaClusterSchedulingPolicySpread

template <int PPT, int CTAS, int TOTAL>
static int launch_cluster_c(int b, int n, int m, int bs, int lg, const float *xyz, float *temp, int *idx, cudaStream_t st,
                            bool probe_only) {
    // > half of the SM's shared memory: ONE CTA per SM, so that co-resident clusters never share an SM (their
    // polling loops would steal issue slots from each other's latency chain) and the occupancy query below
    // counts exclusive placements
    size_t smem = (size_t)3 * n * sizeof(float);
    if (smem < 120 * 1024) smem = 120 * 1024;
    auto kern = fps_cluster_kernel<PPT, CTAS, TOTAL>;
    KDPC_ENSURE_SMEM(kern, 3 * 16384 * (int)sizeof(float));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)b * CTAS);
    cfg.blockDim = dim3(TOTAL / CTAS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;
    attr[1].val.clusterSchedulingPolicyPreference = cudaClusterSchedulingPolicySpread;
    cfg.attrs = attr;
    cfg.numAttrs = fps_cluster_spread ? 2 : 1;
    if (probe_only) {                                          // how many clusters of this shape fit at once?
        int clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
        return clusters;
    }
    return (int)cudaLaunchKernelEx(&cfg, kern, n, m, bs, lg, xyz, temp, idx);
}

// All clouds must be co-resident (a cluster that waits for SMs just serialises the latency chain): 8 CTAs per
// cloud when the batch fits (16 clusters of 8 need every GPC), else 4 CTAs per cloud, else one CTA per cloud.
template <int PPT>
static int launch_cluster(int b, int n, int m, int bs, int lg, const float *xyz, float *temp, int *idx, cudaStream_t st) {
    static int fit8[4][64], fit4[4][64];                       // [log2 PPT][device], 0 = not probed yet
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = PPT == 1 ? 0 : (PPT == 2 ? 1 : (PPT == 4 ? 2 : 3));
    if (dev < 0 || dev >= 64) return -100;
    if (fit8[slot][dev] == 0) {
        fit8[slot][dev] = 1 + launch_cluster_c<PPT, 8, FPSC_TOTAL>(1, n, m, bs, lg, xyz, temp, idx, st, true);
        fit4[slot][dev] = 1 + launch_cluster_c<PPT, 4, FPSC_TOTAL>(1, n, m, bs, lg, xyz, temp, idx, st, true);
    }
    if (b <= fit8[slot][dev] - 1) return launch_cluster_c<PPT, 8, FPSC_TOTAL>(b, n, m, bs, lg, xyz, temp, idx, st, false);
    if (b <= fit4[slot][dev] - 1 && b <= 24)                   // (measured: beyond ~24 clouds the one-CTA kernel wins)
        return launch_cluster_c<PPT, 4, FPSC_TOTAL>(b, n, m, bs, lg, xyz, temp, idx, st, false);
    return -100;                                               // caller falls back to the one-CTA kernel
}

// Forced variant (measurements: tools/bench_fps.py): ctas in {2,4,8}, total threads per cloud in {1024, 2048}.
// Returns -100 when the shape does not divide.
template <int CTAS, int TOTAL>
static int launch_cluster_forced(int per, int b, int n, int m, int bs, int lg, const float *xyz, float *temp, int *idx, cudaStream_t st) {
    if (per == 1) return launch_cluster_c<1, CTAS, TOTAL>(b, n, m, bs, lg, xyz, temp, idx, st, false);
    if (per == 2) return launch_cluster_c<2, CTAS, TOTAL>(b, n, m, bs, lg, xyz, temp, idx, st, false);
    if (per == 4) return launch_cluster_c<4, CTAS, TOTAL>(b, n, m, bs, lg, xyz, temp, idx, st, false);
    if (per == 8) return launch_cluster_c<8, CTAS, TOTAL>(b, n, m, bs, lg, xyz, temp, idx, st, false);
    if (per == 16) return launch_cluster_c<16, CTAS, TOTAL>(b, n, m, bs, lg, xyz, temp, idx, st, false);
    return -100;
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

static int kdpc_fps_use_cluster = 1;
/* test / measurement hook: 0 = always the single-CTA kernel, 1 = automatic (default);
   otherwise a forced cluster variant: on = spread * 1000000 + ctas * 10000 + threads_per_cloud (e.g. 82048, 41024, 1082048) */
KDPC_API void kdpc_fps_set_cluster(int on) { kdpc_fps_use_cluster = on; }
/* how many clusters of (ctas, threads per cloud) fit on the device at once for an n-point cloud (0: none / bad shape) */
KDPC_API int kdpc_fps_cluster_capacity(int ctas, int total, int n) {
    using namespace kdpc;
    const int bs = ref_block_size(n);
    int lg = 0;
    while ((1 << lg) < bs) ++lg;
    if (ctas == 8 && total == 2048) return launch_cluster_c<4, 8, 2048>(1, n, 1, bs, lg, nullptr, nullptr, nullptr, 0, true);
    if (ctas == 4 && total == 2048) return launch_cluster_c<4, 4, 2048>(1, n, 1, bs, lg, nullptr, nullptr, nullptr, 0, true);
    if (ctas == 8 && total == 1024) return launch_cluster_c<8, 8, 1024>(1, n, 1, bs, lg, nullptr, nullptr, nullptr, 0, true);
    if (ctas == 4 && total == 1024) return launch_cluster_c<8, 4, 1024>(1, n, 1, bs, lg, nullptr, nullptr, nullptr, 0, true);
    if (ctas == 2 && total == 1024) return launch_cluster_c<8, 2, 1024>(1, n, 1, bs, lg, nullptr, nullptr, nullptr, 0, true);
    return 0;
}

KDPC_API int kdpc_fps(int b, int n, int m, const float *xyz, float *temp, int *idx, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(xyz && idx && b > 0 && n > 0);
    if (m <= 0) return KDPC_OK;                       // sampling_gpu.cu:100
    using namespace kdpc;
    cudaStream_t st = to_stream(stream);
    const int bs = ref_block_size(n);
    int lg = 0;
    while ((1 << lg) < bs) ++lg;
    const int ppt = (n + bs - 1) / bs;
    // clouds of >= 2048 points: a cluster of 8 CTAs per cloud (2048 threads = 2048/bs threads per reference thread)
    if (kdpc_fps_use_cluster > 1) {                   // forced variant (measurements)
        const int spread = kdpc_fps_use_cluster / 1000000, ctas = (kdpc_fps_use_cluster / 10000) % 100,
                  tot = kdpc_fps_use_cluster % 10000;
        fps_cluster_spread = spread;
        int rc = -100;
        if (n <= 16384 && tot >= bs && (ppt * bs) % tot == 0) {
            const int per = ppt * bs / tot;
            if (ctas == 8 && tot == 2048) rc = launch_cluster_forced<8, 2048>(per, b, n, m, bs, lg, xyz, temp, idx, st);
            if (ctas == 4 && tot == 2048) rc = launch_cluster_forced<4, 2048>(per, b, n, m, bs, lg, xyz, temp, idx, st);
            if (ctas == 2 && tot == 2048) rc = launch_cluster_forced<2, 2048>(per, b, n, m, bs, lg, xyz, temp, idx, st);
            if (ctas == 8 && tot == 1024) rc = launch_cluster_forced<8, 1024>(per, b, n, m, bs, lg, xyz, temp, idx, st);
            if (ctas == 4 && tot == 1024) rc = launch_cluster_forced<4, 1024>(per, b, n, m, bs, lg, xyz, temp, idx, st);
            if (ctas == 2 && tot == 1024) rc = launch_cluster_forced<2, 1024>(per, b, n, m, bs, lg, xyz, temp, idx, st);
        }
        fps_cluster_spread = 0;
        if (rc != -100) return rc;
    }
    const int total = FPSC_TOTAL;
    if (kdpc_fps_use_cluster && n >= 2 * total && n <= 16384 && (ppt * bs) % total == 0) {    // (n = 2048: one CTA is faster)
        const int per = ppt * bs / total;
        int rc = -100;
        if (per == 1) rc = launch_cluster<1>(b, n, m, bs, lg, xyz, temp, idx, st);
        if (per == 2) rc = launch_cluster<2>(b, n, m, bs, lg, xyz, temp, idx, st);
        if (per == 4) rc = launch_cluster<4>(b, n, m, bs, lg, xyz, temp, idx, st);
        if (per == 8) rc = launch_cluster<8>(b, n, m, bs, lg, xyz, temp, idx, st);
        if (rc != -100) return rc;
    }
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
