// Exact k-nearest neighbours by best-first search over spatially sorted tiles (sm_100a).
//
// Same contract and bit-identical results as the brute-force kernel in knn.cu (ascending
// (distance, index) with the reference's  -2 q.c + |q|^2 + |c|^2  rounding sequence,
// pointconv_util.py:73-107), but the S x N distance matrix is never even *computed* in full:
//
//   1. spatial_sort_kernel (one CTA per cloud): Morton-orders the cloud (register/shuffle/shared-memory
//      bitonic sort of 32-bit (18-bit code, index) keys), writes the points as float4 (x, y, z, |p|^2) in that
//      order plus the original indices, and one bounding box per tile of 64 consecutive points.
//   2. knn_bf_kernel: one WARP per query.  The warp ranks the candidate tiles by a lower bound of the
//      distance between the query and the tile box, visits them nearest-first (64 candidates = two
//      coalesced 32-wide steps, software-prefetched one tile ahead) and STOPS at the first tile whose
//      bound exceeds the current K-th distance.  The K best are kept distributed over the lanes (lane j
//      = j-th smallest), so an insertion is a ballot + shuffle-up.  At 8192 uniformly distributed
//      points a query visits ~5 (K = 3) to ~10 (K = 32) of the 128 tiles.
//
// Exactness.  The bound is conservative w.r.t. the rounding of the expansion formula: for any
// candidate c in a tile,  d_fp32(q, c) >= |q-c|^2 - 10u(|q|^2+|c|^2)  (u = 2^-24; derivation in
// DESIGN.md), and the tile is skipped only if  boxdist^2 - 4e-6 (max|q|^2 + max|c|^2) > tau_max,
// so no candidate that could enter the list (d <= tau) is ever skipped.  Ties are resolved
// by the explicit (distance, index) order, independent of the visiting order.
#include "common.cuh"

namespace kdpc {

constexpr int BF_TILE = 64;              // candidates per tile
constexpr int BF_MAX_N = 16384;          // one CTA sorts a whole cloud in shared memory
constexpr int SORT_THREADS = 1024;
constexpr float BF_MARGIN = 4e-6f;

// ---- 1. spatial sort -------------------------------------------------------------------------
// ws layout per cloud b (all 16-byte aligned): sorted4 [n] float4 | boxes [ntiles] 2 x float4 | sidx [n] int | inv [n] int
struct SortedCloud {
    float4 *p4;
    float4 *boxes;
    int *sidx;      // sorted position -> original index
    int *inv;       // original index -> sorted position
};
__host__ __device__ static inline size_t sorted_idx_bytes(int n) { return (((size_t)n * 4 + 15) / 16) * 16; }
static inline size_t sorted_cloud_bytes(int n) {
    const size_t nt = (size_t)(n + BF_TILE - 1) / BF_TILE;
    return (size_t)n * 16 + nt * 32 + 2 * sorted_idx_bytes(n);
}
__host__ __device__ static inline SortedCloud sorted_cloud_at(void *ws, int b, int n) {
    const size_t nt = (size_t)(n + BF_TILE - 1) / BF_TILE;
    const size_t per = (size_t)n * 16 + nt * 32 + 2 * sorted_idx_bytes(n);
    unsigned char *base = reinterpret_cast<unsigned char *>(ws) + per * (size_t)b;
    SortedCloud c;
    c.p4 = reinterpret_cast<float4 *>(base);
    c.boxes = reinterpret_cast<float4 *>(base + (size_t)n * 16);
    c.sidx = reinterpret_cast<int *>(base + (size_t)n * 16 + nt * 32);
    c.inv = reinterpret_cast<int *>(base + (size_t)n * 16 + nt * 32 + sorted_idx_bytes(n));
    return c;
}

// Sort key: 18-bit Morton code (6 bits per axis, cubic cells) << 14 | point index (n <= 16384): 32-bit keys,
// unique, so the order is deterministic.  Bitonic network with E consecutive elements per thread:
// strides < E are exchanged in registers, strides < 32E by warp shuffles, only the longer ones through
// shared memory (15 of the 91 sub-stages at n = 8192).
__device__ __forceinline__ unsigned expand6(unsigned v) {      // 6 bits -> every third bit
    v = (v | (v << 8)) & 0x0000300Fu;
    v = (v | (v << 4)) & 0x000030C3u;
    v = (v | (v << 2)) & 0x00009249u;
    return v;
}

template <int E>
__global__ void __launch_bounds__(SORT_THREADS)
spatial_sort_kernel(int n, int n2 /* pow2 >= n, = E * blockDim.x */, const float *__restrict__ xyz, void *__restrict__ ws) {
    extern __shared__ __align__(16) unsigned skeys[];                      // n2 keys
    __shared__ float red[6][32];
    __shared__ float bb[6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int b = blockIdx.x;
    const float *p = xyz + (size_t)b * n * 3;

    // cloud bounding box
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < n; i += nthreads) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = p[i * 3 + c];
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if (lane == 0) { red[c][warp] = lo[c]; red[3 + c][warp] = hi[c]; }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float l = lane < nwarps ? red[c][lane] : INFINITY, h = lane < nwarps ? red[3 + c][lane] : -INFINITY;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
                h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
            }
            if (lane == 0) { bb[c] = l; bb[3 + c] = h; }
        }
    }
    __syncthreads();
    const float ext = fmaxf(fmaxf(bb[3] - bb[0], bb[4] - bb[1]), fmaxf(bb[5] - bb[2], 1e-30f));
    const float scale = 63.f / ext;                                        // cubic cells

    unsigned key[E];                                                       // elements tid*E .. tid*E+E-1
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int i = tid * E + r;
        key[r] = 0xffffffffu;
        if (i < n) {
            const unsigned cx = (unsigned)fminf(fmaxf((p[i * 3 + 0] - bb[0]) * scale, 0.f), 63.f);
            const unsigned cy = (unsigned)fminf(fmaxf((p[i * 3 + 1] - bb[1]) * scale, 0.f), 63.f);
            const unsigned cz = (unsigned)fminf(fmaxf((p[i * 3 + 2] - bb[2]) * scale, 0.f), 63.f);
            const unsigned code = (expand6(cx) << 2) | (expand6(cy) << 1) | expand6(cz);
            key[r] = (code << 14) | (unsigned)i;
        }
    }
    // ascending bitonic sort
    for (int kk = 2; kk <= n2; kk <<= 1) {
        for (int j = kk >> 1; j >= E; j >>= 1) {
            if (j >= 32 * E) {                                             // partner in another warp
                __syncthreads();
#pragma unroll
                for (int r = 0; r < E; ++r) skeys[tid * E + r] = key[r];
                __syncthreads();
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const int i = tid * E + r;
                    const unsigned c = skeys[i ^ j];
                    const bool keep_min = ((i & j) == 0) == ((i & kk) == 0);
                    key[r] = keep_min ? min(key[r], c) : max(key[r], c);
                }
            } else {                                                       // partner in another lane, same slot
                const bool keep_min = (((tid * E) & j) == 0) == (((tid * E) & kk) == 0);   // j, kk >= E: slot bits irrelevant
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const unsigned c = __shfl_xor_sync(0xffffffffu, key[r], j / E);
                    key[r] = keep_min ? min(key[r], c) : max(key[r], c);
                }
            }
        }
#pragma unroll
        for (int jj = E >> 1; jj > 0; jj >>= 1) {                          // partner in this thread: static register pairs
            if (jj < kk) {
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    if ((r & jj) == 0) {
                        const bool up = (((tid * E + r) & kk) == 0);
                        const unsigned a = key[r], c = key[r | jj];
                        const bool sw = (a > c) == up;
                        key[r] = sw ? c : a;
                        key[r | jj] = sw ? a : c;
                    }
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < E; ++r) skeys[tid * E + r] = key[r];
    __syncthreads();

    SortedCloud out = sorted_cloud_at(ws, b, n);
    for (int i = tid; i < n; i += nthreads) {
        const int src = (int)(skeys[i] & 0x3fffu);
        const float x = p[src * 3 + 0], y = p[src * 3 + 1], z = p[src * 3 + 2];
        out.p4[i] = make_float4(x, y, z, sq_norm3(x, y, z));
        out.sidx[i] = src;
        out.inv[src] = i;
    }
    // tile boxes: one warp per tile
    const int ntiles = (n + BF_TILE - 1) / BF_TILE;
    for (int t = warp; t < ntiles; t += nwarps) {
        float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY}, cc = 0.f;
#pragma unroll
        for (int e = 0; e < BF_TILE / 32; ++e) {
            const int i = t * BF_TILE + e * 32 + lane;
            if (i < n) {
                const int src = (int)(skeys[i] & 0x3fffu);
                const float x = p[src * 3 + 0], y = p[src * 3 + 1], z = p[src * 3 + 2];
                l[0] = fminf(l[0], x); h[0] = fmaxf(h[0], x);
                l[1] = fminf(l[1], y); h[1] = fmaxf(h[1], y);
                l[2] = fminf(l[2], z); h[2] = fmaxf(h[2], z);
                cc = fmaxf(cc, sq_norm3(x, y, z));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                l[c] = fminf(l[c], __shfl_xor_sync(0xffffffffu, l[c], o));
                h[c] = fmaxf(h[c], __shfl_xor_sync(0xffffffffu, h[c], o));
            }
            cc = fmaxf(cc, __shfl_xor_sync(0xffffffffu, cc, o));
        }
        if (lane == 0) {
            out.boxes[2 * t] = make_float4(l[0], l[1], l[2], cc);
            out.boxes[2 * t + 1] = make_float4(h[0], h[1], h[2], 0.f);
        }
    }
}

// ---- 1b. re-use of a spatial order ---------------------------------------------------------------
// A WARPED cloud (PointWarping: xyz2 - interp(flow), xyz1 + flow; reference pointconv_util.py:2114-2142) is a smooth
// displacement of a cloud whose Morton order is already known.  Its sorted representation is built from the PARENT's
// order instead of a second bitonic sort (59 us per 8192-point batch on one CTA per cloud): points are gathered in the
// parent's order and the tile boxes are recomputed from the NEW coordinates.  The search needs only valid boxes, not a
// good order, so results stay bit-identical; a less coherent displacement just makes the boxes looser (more tiles visited).
// One warp per tile of 64 points.
__global__ void __launch_bounds__(256)
spatial_reorder_kernel(int n, const float *__restrict__ xyz, const void *__restrict__ parent_ws, void *__restrict__ ws) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int ntiles = (n + BF_TILE - 1) / BF_TILE;
    if (t >= ntiles) return;
    const SortedCloud par = sorted_cloud_at(const_cast<void *>(parent_ws), b, n);
    SortedCloud out = sorted_cloud_at(ws, b, n);
    const float *p = xyz + (size_t)b * n * 3;
    float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY}, cc = 0.f;
#pragma unroll
    for (int e = 0; e < BF_TILE / 32; ++e) {
        const int i = t * BF_TILE + e * 32 + lane;
        if (i < n) {
            const int src = par.sidx[i];
            const float x = p[src * 3 + 0], y = p[src * 3 + 1], z = p[src * 3 + 2];
            const float nn = sq_norm3(x, y, z);
            out.p4[i] = make_float4(x, y, z, nn);
            out.sidx[i] = src;
            out.inv[src] = i;
            l[0] = fminf(l[0], x); h[0] = fmaxf(h[0], x);
            l[1] = fminf(l[1], y); h[1] = fmaxf(h[1], y);
            l[2] = fminf(l[2], z); h[2] = fmaxf(h[2], z);
            cc = fmaxf(cc, nn);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            l[c] = fminf(l[c], __shfl_xor_sync(0xffffffffu, l[c], o));
            h[c] = fmaxf(h[c], __shfl_xor_sync(0xffffffffu, h[c], o));
        }
        cc = fmaxf(cc, __shfl_xor_sync(0xffffffffu, cc, o));
    }
    if (lane == 0) {
        out.boxes[2 * t] = make_float4(l[0], l[1], l[2], cc);
        out.boxes[2 * t + 1] = make_float4(h[0], h[1], h[2], 0.f);
    }
}

// ---- 2. best-first search, one WARP per query ----------------------------------------------------
// A list entry is ONE 64-bit key: (order-preserving bits of the fp32 distance) << 32 | candidate index, so that the
// (distance, index) order is a plain unsigned compare (2 instructions, branch-free) and a shuffle moves both.
__device__ __forceinline__ unsigned long long make_key(float d, int i) {
    const unsigned u = __float_as_uint(d);
    const unsigned o = (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // monotone: -x < +0 < +x < +inf
    return ((unsigned long long)o << 32) | (unsigned)i;
}
__device__ __forceinline__ float key_dist(unsigned long long k) {
    const unsigned o = (unsigned)(k >> 32);
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ int key_idx(unsigned long long k) { return (int)(unsigned)(k & 0xffffffffull); }
constexpr unsigned long long BF_EMPTY = 0xff8000007fffffffull;          // (+inf, INT_MAX)

constexpr int BF_QPW = 8;                // max consecutive (Morton-adjacent) queries per warp
constexpr int BF_CTA_WARPS = 8;
constexpr unsigned BF_NONE = 0xffffffffu;
constexpr int BF_FEW = 5;                // flush of up to this many survivors: one by one
constexpr int BF_FLUSH = 24;             // merge the buffered survivors when this many are waiting (checked per tile)
constexpr int BF_BUF = BF_FLUSH + 64;    // a tile adds at most 64

// The K best of a query live DISTRIBUTED over the warp: lane j holds the j-th smallest (distance, index).
// Candidates are tested 32 at a time (one per lane); those that pass the current K-th distance are appended
// to a small per-warp buffer (ballot + popc) and merged into the list two dozen at a time by a sorting
// network of shuffles (sort the batch, pair it reversed with the list, half-cleaners), or one by one (ballot +
// shuffle-up) when only a handful are waiting; no divergence, no per-thread sorted list.
// Tiles are 1 KB coalesced reads that stay in L1/L2.
// Each lane ranks KEYS tiles (tile = e*32 + lane) by the query's own lower bound; the warp pops the
// nearest remaining tile with one redux.min.
template <int MODE, int KEYS>
__global__ void __launch_bounds__(BF_CTA_WARPS * 32)
knn_bf_kernel(int s, int n, int k, int qpw, const void *__restrict__ qws, const void *__restrict__ cws,
              int *__restrict__ idx32, long long *__restrict__ idx64, float *__restrict__ dist_out) {
    __shared__ unsigned long long buf[BF_CTA_WARPS][BF_BUF];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const SortedCloud Q = sorted_cloud_at(const_cast<void *>(qws), b, s);
    const SortedCloud C = sorted_cloud_at(const_cast<void *>(cws), b, n);
    const int ntiles = (n + BF_TILE - 1) / BF_TILE;
    const int qbeg = (blockIdx.x * BF_CTA_WARPS + warp) * qpw;
    const int qend = min(s, qbeg + qpw);

    unsigned long long key = BF_EMPTY;                      // the list survives the loop iteration: it seeds the next query
#pragma unroll 1
    for (int qpos = qbeg; qpos < qend; ++qpos) {
        const float4 q = __ldg(Q.p4 + qpos);
        const int qorig = __ldg(Q.sidx + qpos);

        // lower bound of the fp32 distance to every point of a tile (rounded DOWN to 24 bits) | tile id
        unsigned keys[KEYS];
#pragma unroll
        for (int e = 0; e < KEYS; ++e) {
            const int t = e * 32 + lane;
            keys[e] = BF_NONE;
            if (t < ntiles) {
                const float4 bl = __ldg(C.boxes + 2 * t), bh = __ldg(C.boxes + 2 * t + 1);
                const float dx = fmaxf(0.f, fmaxf(bl.x - q.x, q.x - bh.x));
                const float dy = fmaxf(0.f, fmaxf(bl.y - q.y, q.y - bh.y));
                const float dz = fmaxf(0.f, fmaxf(bl.z - q.z, q.z - bh.z));
                const float lb = fmaf(dx, dx, fmaf(dy, dy, dz * dz)) - BF_MARGIN * (q.w + bl.w);
                keys[e] = lb > 0.f ? ((__float_as_uint(lb) & 0xffffff00u) | (unsigned)t) : (unsigned)t;
            }
        }
        auto pop_min = [&]() -> unsigned {                  // warp-uniform smallest remaining key
            unsigned m = keys[0];
#pragma unroll
            for (int e = 1; e < KEYS; ++e) m = min(m, keys[e]);
            m = __reduce_min_sync(0xffffffffu, m);
#pragma unroll
            for (int e = 0; e < KEYS; ++e) keys[e] = keys[e] == m ? BF_NONE : keys[e];
            return m;
        };
        float4 pre[2];
        int prei[2];
        auto prefetch = [&](unsigned key) {
            const int t = (int)(key & 0xffu);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = t * BF_TILE + h * 32 + lane;
                const bool v = i < n;
                pre[h] = v ? __ldg(C.p4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                prei[h] = v ? __ldg(C.sidx + i) : 0x7fffffff;
            }
        };

        // Seed: the previous query of this warp is Morton-adjacent, so its 32 best candidates are (almost) this
        // query's too.  The K-th smallest of THIS query's distances to those 32 distinct candidates is a valid upper
        // bound of its K-th neighbour distance, known before a single tile is visited: the tile scan then lets
        // through little more than the true neighbours (instead of everything until the list has warmed up).
        float tau = INFINITY;                               // K-th best so far / its upper bound (warp-uniform, never too small)
        bool first = true;
        if (qpos > qbeg) {
            float sd = INFINITY;
            const int pi = key_idx(key);
            if (pi != 0x7fffffff) {
                const float4 c = __ldg(C.p4 + __ldg(C.inv + pi));
                sd = MODE == 0 ? expansion_dist(q.x, q.y, q.z, q.w, c.x, c.y, c.z, c.w)
                               : direct_dist(q.x - c.x, q.y - c.y, q.z - c.z);
            }
#pragma unroll
            for (int kk = 2; kk <= 32; kk <<= 1) {          // ascending bitonic sort of the 32 values
#pragma unroll
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    const float pd = __shfl_xor_sync(0xffffffffu, sd, j);
                    sd = (((lane & j) == 0) == ((lane & kk) == 0)) ? fminf(sd, pd) : fmaxf(sd, pd);
                }
            }
            tau = __shfl_sync(0xffffffffu, sd, k - 1);
            first = false;
        }
        key = BF_EMPTY;                                     // lane j: j-th best so far
        int nb = 0;                                         // survivors waiting in the warp's buffer

        // ascending bitonic sort of one key per lane
        auto sort32 = [&](unsigned long long &x) {
#pragma unroll
            for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, j);
                    const bool keep_min = ((lane & j) == 0) == ((lane & kk) == 0);
                    x = (keep_min == (y < x)) ? y : x;
                }
            }
        };
        // merge up to 32 buffered survivors (one per lane) into the list: sort them, pair the list with the REVERSED
        // batch (lane-wise minimum = the 32 smallest of the 64, as a bitonic sequence), 5 half-cleaner steps
        auto merge_batch = [&](int base) {
            unsigned long long nk = base + lane < nb ? buf[warp][base + lane] : BF_EMPTY;
            sort32(nk);
            const unsigned long long rk = __shfl_sync(0xffffffffu, nk, 31 - lane);
            key = rk < key ? rk : key;
#pragma unroll
            for (int j = 16; j > 0; j >>= 1) {
                const unsigned long long y = __shfl_xor_sync(0xffffffffu, key, j);
                key = (((lane & j) == 0) == (y < key)) ? y : key;
            }
        };
        // one survivor: lanes whose entry comes after it shift up by one; it enters iff lane k-1 is one of them
        auto insert_one = [&](unsigned long long x) {
            const bool before = x < key;
            const unsigned bm = __ballot_sync(0xffffffffu, before);
            if (bm & (1u << (k - 1))) {                     // warp-uniform
                const int pos = __ffs(bm) - 1;
                const unsigned long long up = __shfl_up_sync(0xffffffffu, key, 1);
                if (before) key = lane == pos ? x : up;
            }
        };
        auto kth = [&]() { return key_dist(__shfl_sync(0xffffffffu, key, k - 1)); };
        auto flush = [&]() {
            __syncwarp();
            if (nb <= BF_FEW) {                             // a handful: one-by-one is cheaper than the sorting network
                for (int e = 0; e < nb; ++e) insert_one(buf[warp][e]);
            } else {
                for (int base = 0; base < nb; base += 32) merge_batch(base);
            }
            __syncwarp();
            nb = 0;
            tau = fminf(tau, kth());
        };

        unsigned cur = pop_min();
        if (cur != BF_NONE) prefetch(cur);
        while (cur != BF_NONE) {
            // bound > 0 and beyond the K-th best: so is every remaining tile (tau may be stale, i.e. too LARGE: safe)
            if (cur >= 256u && __uint_as_float(cur & 0xffffff00u) > tau) break;
            const float4 c0 = pre[0], c1 = pre[1];
            const int ci0 = prei[0], ci1 = prei[1];
            cur = pop_min();
            if (cur != BF_NONE) prefetch(cur);              // next tile in flight while this one is scanned
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 c = h ? c1 : c0;
                float d = MODE == 0 ? expansion_dist(q.x, q.y, q.z, q.w, c.x, c.y, c.z, c.w)
                                    : direct_dist(q.x - c.x, q.y - c.y, q.z - c.z);
                const int i = h ? ci1 : ci0;
                if (i == 0x7fffffff) d = INFINITY;          // slot beyond the cloud
                if (first) {                                // the first 32 candidates, sorted, ARE the initial list
                    first = false;
                    key = make_key(d, i);
                    sort32(key);
                    tau = kth();
                } else {
                    // survivors of the (possibly stale) K-th distance are only APPENDED to the warp's buffer ...
                    const bool pass = d <= tau;
                    const unsigned mask = __ballot_sync(0xffffffffu, pass);
                    if (pass) buf[warp][nb + __popc(mask & ((1u << lane) - 1u))] = make_key(d, i);
                    nb += __popc(mask);
                }
            }
            if (nb >= BF_FLUSH) flush();                    // ... and merged 24+ at a time by a sorting network
        }
        if (nb > 0) flush();
        if (lane < k) {
            const size_t o = ((size_t)b * s + qorig) * k + lane;
            if (idx32) idx32[o] = key_idx(key);
            if (idx64) idx64[o] = key_idx(key);
            if (dist_out) dist_out[o] = key_dist(key);
        }
    }
}

// ---- 2b. K <= 4 (3-NN of the warping / upsampling layers): one THREAD per query ----------------------------------
// With three neighbours a whole warp per query spends its time on list upkeep and tile ranking, not on distances.  Here
// a warp owns 32 Morton-consecutive queries and walks the candidate tiles ONCE for all of them: tiles are ranked by the
// lower bound between the tile box and the WARP's query box (lanes hold KEYS tile keys each, redux.min pops the nearest),
// a tile's 64 points are staged in shared memory and every lane tests all of them against its own query (broadcast
// LDS.128, the K best as sorted 64-bit (distance, index) keys in registers), and the walk stops at the first tile whose
// bound exceeds the LARGEST K-th distance of the 32 lanes.  More tiles per query than the per-query walk (~14 instead of
// ~6), ~1/8 of the instructions per candidate; same (distance, index) order, so the results are bit-identical.
constexpr int FEW_WARPS = 4, FEW_KMAX = 4;

template <int MODE, int KEYS>
__global__ void __launch_bounds__(FEW_WARPS * 32)
knn_few_kernel(int s, int n, int k, const void *__restrict__ qws, const void *__restrict__ cws, int *__restrict__ idx32,
               long long *__restrict__ idx64, float *__restrict__ dist_out) {
    __shared__ float4 tp[FEW_WARPS][BF_TILE];
    __shared__ int ti[FEW_WARPS][BF_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const SortedCloud Q = sorted_cloud_at(const_cast<void *>(qws), b, s);
    const SortedCloud C = sorted_cloud_at(const_cast<void *>(cws), b, n);
    const int ntiles = (n + BF_TILE - 1) / BF_TILE;
    const int qpos = (blockIdx.x * FEW_WARPS + warp) * 32 + lane;
    if ((blockIdx.x * FEW_WARPS + warp) * 32 >= s) return;             // whole warp beyond the cloud
    const bool valid = qpos < s;
    const float4 q = __ldg(Q.p4 + (valid ? qpos : s - 1));             // (idle lanes repeat the last query: the box stays tight)
    // the warp's query box and largest |q|^2
    float lo[3] = {q.x, q.y, q.z}, hi[3] = {q.x, q.y, q.z}, qqmax = q.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        qqmax = fmaxf(qqmax, __shfl_xor_sync(0xffffffffu, qqmax, o));
    }
    unsigned keys[KEYS];
#pragma unroll
    for (int e = 0; e < KEYS; ++e) {
        const int t = e * 32 + lane;
        keys[e] = BF_NONE;
        if (t < ntiles) {
            const float4 bl = __ldg(C.boxes + 2 * t), bh = __ldg(C.boxes + 2 * t + 1);
            const float dx = fmaxf(0.f, fmaxf(bl.x - hi[0], lo[0] - bh.x));
            const float dy = fmaxf(0.f, fmaxf(bl.y - hi[1], lo[1] - bh.y));
            const float dz = fmaxf(0.f, fmaxf(bl.z - hi[2], lo[2] - bh.z));
            const float lb = fmaf(dx, dx, fmaf(dy, dy, dz * dz)) - BF_MARGIN * (qqmax + bl.w);
            keys[e] = lb > 0.f ? ((__float_as_uint(lb) & 0xffffff00u) | (unsigned)t) : (unsigned)t;
        }
    }
    unsigned long long best[FEW_KMAX];
#pragma unroll
    for (int j = 0; j < FEW_KMAX; ++j) best[j] = BF_EMPTY;
    float tau = INFINITY;                                               // this lane's K-th distance so far
    while (true) {
        unsigned cur = keys[0];
#pragma unroll
        for (int e = 1; e < KEYS; ++e) cur = min(cur, keys[e]);
        cur = __reduce_min_sync(0xffffffffu, cur);
        if (cur == BF_NONE) break;
#pragma unroll
        for (int e = 0; e < KEYS; ++e) keys[e] = keys[e] == cur ? BF_NONE : keys[e];
        // largest K-th distance of the warp (taus are >= -tiny or +inf: compare through the order-preserving map)
        const unsigned tmax = __reduce_max_sync(0xffffffffu, (unsigned)(make_key(tau, 0) >> 32));
        if (cur >= 256u && __uint_as_float(cur & 0xffffff00u) > key_dist((unsigned long long)tmax << 32)) break;
        const int t = (int)(cur & 0xffu);
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = t * BF_TILE + h * 32 + lane;
            const bool v = i < n;
            // slots beyond the cloud: a point at infinity (distance +inf in either formula), index INT_MAX
            tp[warp][h * 32 + lane] = v ? __ldg(C.p4 + i) : (MODE == 0 ? make_float4(0.f, 0.f, 0.f, INFINITY) : make_float4(INFINITY, 0.f, 0.f, 0.f));
            ti[warp][h * 32 + lane] = v ? __ldg(C.sidx + i) : 0x7fffffff;
        }
        __syncwarp();
#pragma unroll 8
        for (int j = 0; j < BF_TILE; ++j) {
            const float4 c = tp[warp][j];
            const float d = MODE == 0 ? expansion_dist(q.x, q.y, q.z, q.w, c.x, c.y, c.z, c.w)
                                      : direct_dist(q.x - c.x, q.y - c.y, q.z - c.z);
            if (d <= tau) {                                             // (ties on the distance: the index decides below)
                unsigned long long x = make_key(d, ti[warp][j]);
#pragma unroll
                for (int r = 0; r < FEW_KMAX; ++r) {
                    if (r < k) {
                        const unsigned long long lo64 = x < best[r] ? x : best[r];
                        x = x < best[r] ? best[r] : x;
                        best[r] = lo64;
                    }
                }
                tau = key_dist(best[k - 1]);
            }
        }
    }
    if (valid) {
        const int qorig = __ldg(Q.sidx + qpos);
        const size_t o = ((size_t)b * s + qorig) * k;
#pragma unroll
        for (int r = 0; r < FEW_KMAX; ++r) {
            if (r < k) {
                if (idx32) idx32[o + r] = key_idx(best[r]);
                if (idx64) idx64[o + r] = key_idx(best[r]);
                if (dist_out) dist_out[o + r] = key_dist(best[r]);
            }
        }
    }
}

static inline int next_pow2(int n) {
    int p = 64;
    while (p < n) p <<= 1;
    return p;
}

static int launch_sort(int b, int n, const float *xyz, void *ws, cudaStream_t st) {
    const int n2 = next_pow2(n);
    const int threads = n2 < SORT_THREADS ? n2 : SORT_THREADS;
    const size_t smem = (size_t)n2 * 4;
    switch (n2 / threads) {
        case 1: spatial_sort_kernel<1><<<b, threads, smem, st>>>(n, n2, xyz, ws); break;
        case 2: spatial_sort_kernel<2><<<b, threads, smem, st>>>(n, n2, xyz, ws); break;
        case 4: spatial_sort_kernel<4><<<b, threads, smem, st>>>(n, n2, xyz, ws); break;
        case 8: spatial_sort_kernel<8><<<b, threads, smem, st>>>(n, n2, xyz, ws); break;
        case 16:
            KDPC_ENSURE_SMEM(spatial_sort_kernel<16>, BF_MAX_N * 4);
            spatial_sort_kernel<16><<<b, threads, smem, st>>>(n, n2, xyz, ws);
            break;
        default: return KDPC_EUNSUPPORTED;
    }
    return (int)cudaGetLastError();
}

static int kdpc_knn_few = 0;           // (1: K <= 4 by knn_few_kernel - measured 4x SLOWER than one warp per query, kept for the record; tests compare the two)

template <int MODE>
static int launch_bf(int b, int s, int n, int k, const void *qws, const void *cws, int *idx32, long long *idx64,
                     float *dist, cudaStream_t st) {
    const int ntiles_ = (n + BF_TILE - 1) / BF_TILE;
    if (k <= FEW_KMAX && kdpc_knn_few && (long long)b * s >= 64LL * device_sms()) {
        // few neighbours, enough queries to fill the machine with one thread each: shared tile walk per warp
        dim3 grid((s + FEW_WARPS * 32 - 1) / (FEW_WARPS * 32), b);
#define KDPC_FEW_CASE(KEYS) \
        if (ntiles_ <= 32 * KEYS) { \
            knn_few_kernel<MODE, KEYS><<<grid, FEW_WARPS * 32, 0, st>>>(s, n, k, qws, cws, idx32, idx64, dist); \
            return (int)cudaGetLastError(); }
        KDPC_FEW_CASE(1)
        KDPC_FEW_CASE(2)
        KDPC_FEW_CASE(4)
        KDPC_FEW_CASE(8)
#undef KDPC_FEW_CASE
    }
    // consecutive (Morton-adjacent) queries per warp: up to BF_QPW, fewer when the call is small so that the
    // machine still gets ~48 warps per SM
    int qpw = (int)(((long long)b * s) / (device_sms() * 48));
    qpw = qpw < 1 ? 1 : (qpw > BF_QPW ? BF_QPW : qpw);
    const int per_cta = BF_CTA_WARPS * qpw;
    dim3 grid((s + per_cta - 1) / per_cta, b);
    const int ntiles = (n + BF_TILE - 1) / BF_TILE;
#define KDPC_BF_CASE(KEYS) \
    if (ntiles <= 32 * KEYS) { \
        knn_bf_kernel<MODE, KEYS><<<grid, BF_CTA_WARPS * 32, 0, st>>>(s, n, k, qpw, qws, cws, idx32, idx64, dist); \
        return (int)cudaGetLastError(); }
    KDPC_BF_CASE(1)
    KDPC_BF_CASE(2)
    KDPC_BF_CASE(4)
    KDPC_BF_CASE(8)
#undef KDPC_BF_CASE
    return KDPC_EUNSUPPORTED;
}

}  // namespace kdpc

using namespace kdpc;

/* A/B switch for measurements and tests: 0 = K <= 4 searches also run one warp per query (same results). */
KDPC_API void kdpc_knn_set_few(int on) { kdpc::kdpc_knn_few = on; }

KDPC_API int kdpc_spatial_sort_order_offset(int n) {
    const size_t nt = (size_t)(n + BF_TILE - 1) / BF_TILE;
    return (int)(((size_t)n * 16 + nt * 32) / 4);
}
KDPC_API int kdpc_spatial_sort_order_stride(int n) { return (int)(sorted_cloud_bytes(n) / 4); }

KDPC_API long long kdpc_spatial_sort_bytes(int b, int n) {
    if (b <= 0 || n <= 0) return 0;
    return (long long)sorted_cloud_bytes(n) * b;
}

KDPC_API int kdpc_spatial_sort(int b, int n, const float *xyz, void *ws, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(xyz && ws && b > 0 && n > 0);
    if (n > BF_MAX_N) return KDPC_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(ws) % 16) != 0) return KDPC_EINVAL;
    return launch_sort(b, n, xyz, ws, to_stream(stream));
}

/* Sorted representation of xyz [B,N,3] in the ORDER of an already sorted cloud of the same size (parent_ws from
 * kdpc_spatial_sort / kdpc_spatial_reorder): for displaced copies of a cloud (warping); boxes are recomputed, results of
 * kdpc_knn_sorted are unchanged. */
KDPC_API int kdpc_spatial_reorder(int b, int n, const float *xyz, const void *parent_ws, void *ws, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(xyz && parent_ws && ws && b > 0 && n > 0);
    if (n > BF_MAX_N || b > 65535) return KDPC_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(ws) % 16) != 0 || (reinterpret_cast<uintptr_t>(parent_ws) % 16) != 0) return KDPC_EINVAL;
    const int ntiles = (n + BF_TILE - 1) / BF_TILE;
    dim3 grid((ntiles + 7) / 8, b);
    spatial_reorder_kernel<<<grid, 256, 0, to_stream(stream)>>>(n, xyz, parent_ws, ws);
    KDPC_RETURN_LAST();
}

KDPC_API int kdpc_knn_sorted(int b, int s, int n, int k, int direct, const void *query_sorted, const void *cand_sorted,
                             int *idx32, long long *idx64, float *dist, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(query_sorted && cand_sorted && b > 0 && s > 0 && n > 0 && k > 0);
    if (k > 32 || k > n || b > 65535 || n > BF_MAX_N) return KDPC_EUNSUPPORTED;
    if (direct) return launch_bf<1>(b, s, n, k, query_sorted, cand_sorted, idx32, idx64, dist, to_stream(stream));
    return launch_bf<0>(b, s, n, k, query_sorted, cand_sorted, idx32, idx64, dist, to_stream(stream));
}
