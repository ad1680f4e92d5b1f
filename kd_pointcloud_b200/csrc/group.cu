// Gather / grouping kernels for sm_100a.
//
// Channel-major (pointnet2 API) versions replace gather_points_kernel_fast
// (reference pointnet2/src/sampling_gpu.cu:8-24) and group_points_kernel_fast
// (group_points_gpu.cu:47-66); point-major versions replace the permute+contiguous+op+permute
// adapters index_points_gather / index_points_group and the group / group_query op chains
// (pointconv_util.py:109-182).  All of them are pure data movement => HBM-bound; the design
// goal is one pass, 128-bit accesses, and indices read once.
#include "common.cuh"

namespace kdpc {

// ------------------------------------------------------------------------------------------
// channel-major gather: out[b,c,j] = f[b,c,idx[b,j]].  On the model path C = 3 (new_xyz, GT flow).
__global__ void gather_cm_kernel(int c, int n, int m, const float *__restrict__ f, const int *__restrict__ idx,
                                 float *__restrict__ out) {
    const int b = blockIdx.z, ci = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int src = __ldg(idx + (size_t)b * m + j);
    out[((size_t)b * c + ci) * m + j] = __ldg(f + ((size_t)b * c + ci) * n + src);
}

// channel-major grouping: out[b,c,s,k] = f[b,c,idx[b,s,k]].
// A CTA stages CPB whole channel rows (N floats each) in shared memory, then streams its share of the index list
// for those channels: random reads hit shared memory instead of L2 sectors (a 4-byte random global read moves a
// 32-byte sector) and the output is written with 128-bit stores, four consecutive (s,k) entries per thread.
// CPB = 2 keeps the staging at 64 KB for N = 8192, so three CTAs share an SM and one CTA's staging overlaps the
// others' gathers (CPB = 4 with one resident CTA per SM was 1.5x SLOWER than the reference's direct kernel).
template <int CPB>
__global__ void __launch_bounds__(256)
group_cm_smem_kernel(int c, int n, int sk, int chunk, const float *__restrict__ f, const int *__restrict__ idx,
                     float *__restrict__ out) {
    extern __shared__ __align__(16) float rows[];        // [CPB][n]
    const int b = blockIdx.z, c0 = blockIdx.y * CPB;
    const float *fb = f + ((size_t)b * c + c0) * n;
    const int nch = min(CPB, c - c0);
    const int tot = nch * n;
    if ((reinterpret_cast<uintptr_t>(fb) & 15) == 0 && (tot & 3) == 0) {
        const float4 *f4 = reinterpret_cast<const float4 *>(fb);
        float4 *r4 = reinterpret_cast<float4 *>(rows);
        for (int i = threadIdx.x; i < (tot >> 2); i += blockDim.x) r4[i] = ld_stream_f4(f4 + i);
    } else {
        for (int i = threadIdx.x; i < tot; i += blockDim.x) rows[i] = fb[i];
    }
    __syncthreads();
    const int *ib = idx + (size_t)b * sk;
    float *ob = out + ((size_t)b * c + c0) * sk;
    const int e0 = blockIdx.x * chunk, e1 = min(sk, e0 + chunk);          // chunk % 4 == 0
    const bool vec = (sk & 3) == 0 && ((reinterpret_cast<uintptr_t>(ib) | reinterpret_cast<uintptr_t>(ob)) & 15) == 0;
    if (vec) {
        for (int e = e0 + 4 * threadIdx.x; e < e1; e += 4 * blockDim.x) {
            const int4 src = __ldg(reinterpret_cast<const int4 *>(ib + e));
#pragma unroll
            for (int ci = 0; ci < CPB; ++ci)
                if (ci < nch) {
                    const float *r = rows + ci * n;
                    st_stream_f4(reinterpret_cast<float4 *>(ob + (size_t)ci * sk + e), make_float4(r[src.x], r[src.y], r[src.z], r[src.w]));
                }
        }
    } else {
        for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
            const int src = ib[e];
#pragma unroll
            for (int ci = 0; ci < CPB; ++ci)
                if (ci < nch) ob[(size_t)ci * sk + e] = rows[ci * n + src];
        }
    }
}

__global__ void group_cm_direct_kernel(int c, int n, int sk, const float *__restrict__ f, const int *__restrict__ idx,
                                       float *__restrict__ out) {
    const int b = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= sk) return;
    const int src = idx[(size_t)b * sk + e];
    const float *fb = f + (size_t)b * c * n;
    float *ob = out + (size_t)b * c * sk;
    for (int ci = 0; ci < c; ++ci) ob[(size_t)ci * sk + e] = __ldg(fb + (size_t)ci * n + src);
}

// ------------------------------------------------------------------------------------------
// point-major row gather: out[r,:] = f[b(r), idx[r], :]   (rows r = b*m + j, C floats per row).
// VEC = 4: C % 4 == 0 and both bases 16-byte aligned -> one float4 per thread, consecutive threads
// cover consecutive 16-byte pieces of the same output row (fully coalesced stores, row-contiguous
// 128-bit loads).
// 16-byte variant for < 2^32 pieces: four pieces per thread (a warp covers 4 x 512 contiguous output bytes), the four
// index loads and then the four row loads in flight before the first store, 32-bit index arithmetic, streaming stores
// (the one-piece-per-thread kernel below reached 50 % of the copy peak at B=8, N=8192, K=16, C=64).
__global__ void __launch_bounds__(256)
gather_rows4_kernel(unsigned total, unsigned n, unsigned m, unsigned cvec, const float4 *__restrict__ f,
                    const int *__restrict__ idx, float4 *__restrict__ out) {
    const unsigned base = blockIdx.x * 1024u + threadIdx.x;
    unsigned src[4], cv[4], bb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned e = base + 256u * j;
        const unsigned r = e / cvec;
        cv[j] = e - r * cvec;
        bb[j] = r / m;
        src[j] = e < total ? (unsigned)__ldg(idx + r) : 0u;
    }
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (base + 256u * j < total) v[j] = __ldg(f + ((size_t)bb[j] * n + src[j]) * cvec + cv[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (base + 256u * j < total) st_stream_f4(out + base + 256u * j, v[j]);
}

template <int VEC>
__global__ void gather_rows_kernel(long long total, int n, int m, int cvec, const float *__restrict__ f,
                                   const int *__restrict__ idx, float *__restrict__ out) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long r = e / cvec;
    const int cv = (int)(e - r * cvec);
    const long long b = r / m;
    const int src = __ldg(idx + r);
    const size_t in_off = ((size_t)b * n + src) * (size_t)cvec + cv;
    if (VEC == 4) {
        reinterpret_cast<float4 *>(out)[e] = __ldg(reinterpret_cast<const float4 *>(f) + in_off);
    } else {
        out[e] = __ldg(f + in_off);
    }
}

// ------------------------------------------------------------------------------------------
// fused group + relative xyz + concat (pointconv_util.py:135-182):
//   out[r, 0:3]   = cand_xyz[b, idx[r]] - query_xyz[b, s(r)]
//   out[r, 3:3+D] = feats[b, idx[r], :]
// Row width W = 3 + D is odd-sized (67, 131, ...), so rows are assembled in shared memory and the
// CTA's contiguous output span [r0*W, (r0+R)*W) is then written with aligned 128-bit stores.
constexpr int GC_ROWS = 32;          // rows per CTA (multiple of 4 keeps the span 16-byte aligned)
constexpr int GC_THREADS = 256;

__global__ void __launch_bounds__(GC_THREADS)
group_concat_kernel(long long rows, int n, int s, int k, int d, const float *__restrict__ cand_xyz,
                    const float *__restrict__ query_xyz, const float *__restrict__ feats,
                    const int *__restrict__ idx, float *__restrict__ out) {
    extern __shared__ __align__(16) float stage[];       // [GC_ROWS][W]
    __shared__ int src_s[GC_ROWS];
    const int w = 3 + d;
    const long long r0 = (long long)blockIdx.x * GC_ROWS;
    const int nr = (int)min((long long)GC_ROWS, rows - r0);
    const int tid = threadIdx.x;

    if (tid < nr) {
        const long long r = r0 + tid;
        const long long bs_ = r / k;                      // b*s + s_idx
        const long long b = bs_ / s;
        const int src = idx[r];
        src_s[tid] = src;
        const float *cp = cand_xyz + ((size_t)b * n + src) * 3;
        const float *qp = query_xyz + (size_t)bs_ * 3;
        float *o = stage + tid * w;
        o[0] = cp[0] - qp[0];
        o[1] = cp[1] - qp[1];
        o[2] = cp[2] - qp[2];
    }
    __syncthreads();
    if (d > 0) {
        // rows of one CTA may straddle a batch edge: the batch index is derived per row
        if ((d & 3) == 0) {
            const int dv = d >> 2;
            for (int e = tid; e < nr * dv; e += GC_THREADS) {
                const int rl = e / dv, cv = e - rl * dv;
                const long long bb = (r0 + rl) / ((long long)s * k);
                const float4 v = __ldg(reinterpret_cast<const float4 *>(feats + ((size_t)bb * n + src_s[rl]) * d) + cv);
                float *o = stage + rl * w + 3 + cv * 4;
                o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
            }
        } else {
            for (int e = tid; e < nr * d; e += GC_THREADS) {
                const int rl = e / d, ci = e - rl * d;
                const long long bb = (r0 + rl) / ((long long)s * k);
                stage[rl * w + 3 + ci] = __ldg(feats + ((size_t)bb * n + src_s[rl]) * d + ci);
            }
        }
    }
    __syncthreads();
    const long long o0 = r0 * w;                          // multiple of 4 floats (GC_ROWS % 4 == 0)
    const int tot = nr * w;
    const int tot4 = tot >> 2;
    float4 *o4 = reinterpret_cast<float4 *>(out + o0);
    const float4 *s4 = reinterpret_cast<const float4 *>(stage);
    for (int e = tid; e < tot4; e += GC_THREADS) st_stream_f4(o4 + e, s4[e]);
    for (int e = (tot4 << 2) + tid; e < tot; e += GC_THREADS) out[o0 + e] = stage[e];
}

// Warp-per-row variant: lanes walk the row in steps of 32 floats - coalesced 128-byte loads of the feature row, coalesced
// stores of the output row (which starts at an arbitrary 4-byte offset: W is odd), lanes 0..2 produce the relative
// coordinates.  No shared-memory assembly, no CTA barriers; four rows per warp in flight (eight: slower).  (Measured and rejected: thread =
// one 16-byte piece of the flat output with its four floats located by division - instruction-bound, 180 us against the
// staged kernel's 115 us at B=8, N=8192, K=9, D=128.)
constexpr int GCR = 4;              // rows per warp (independent loads in flight)
__global__ void __launch_bounds__(256)
group_concat_rows_kernel(unsigned rows, unsigned n, unsigned s, unsigned k, unsigned d, const float *__restrict__ cand_xyz,
                         const float *__restrict__ query_xyz, const float *__restrict__ feats, const int *__restrict__ idx,
                         float *__restrict__ out) {
    const unsigned lane = threadIdx.x & 31u, warp = (blockIdx.x * 256u + threadIdx.x) >> 5;
    const unsigned w = 3u + d;
    const unsigned r0 = warp * GCR;
    if (r0 >= rows) return;
    unsigned src[GCR], bs[GCR];
    size_t cand[GCR];
#pragma unroll
    for (int j = 0; j < GCR; ++j) {
        const unsigned r = min(r0 + j, rows - 1u);
        bs[j] = r / k;
        src[j] = (unsigned)__ldg(idx + r);
        cand[j] = (size_t)(bs[j] / s) * n + src[j];
    }
    for (unsigned c0 = 0; c0 < d; c0 += 32u) {
        float v[GCR];
#pragma unroll
        for (int j = 0; j < GCR; ++j) v[j] = c0 + lane < d ? __ldg(feats + cand[j] * d + c0 + lane) : 0.f;
#pragma unroll
        for (int j = 0; j < GCR; ++j)
            if (r0 + j < rows && c0 + lane < d) __stcs(out + (size_t)(r0 + j) * w + 3u + c0 + lane, v[j]);
    }
    if (lane < 3u) {
#pragma unroll
        for (int j = 0; j < GCR; ++j)
            if (r0 + j < rows)
                out[(size_t)(r0 + j) * w + lane] = __ldg(cand_xyz + cand[j] * 3 + lane) - __ldg(query_xyz + (size_t)bs[j] * 3 + lane);
    }
}

// ------------------------------------------------------------------------------------------
// ball query (ball_query_gpu.cu:9-45): first nsample candidates inside the radius, in index
// order, padded with the first hit; rows with no hit are zero.  Candidates are staged in shared
// memory tiles; a warp leaves the scan when all of its queries are full.
constexpr int BQ_THREADS = 128;
constexpr int BQ_TILE = 1024;

__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(int n, int m, float radius, int nsample, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int *__restrict__ idx) {
    __shared__ float tile[BQ_TILE * 3];
    const int b = blockIdx.y;
    const int q = blockIdx.x * BQ_THREADS + threadIdx.x;
    const bool active = q < m;
    const float r2 = __fmul_rn(radius, radius);
    float qx = 0.f, qy = 0.f, qz = 0.f;
    int *o = nullptr;
    if (active) {
        const float *qp = new_xyz + ((size_t)b * m + q) * 3;
        qx = qp[0]; qy = qp[1]; qz = qp[2];
        o = idx + ((size_t)b * m + q) * nsample;
    }
    int cnt = active ? 0 : nsample;
    const float *pb = xyz + (size_t)b * n * 3;
    for (int t0 = 0; t0 < n; t0 += BQ_TILE) {
        const int tn = min(BQ_TILE, n - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * 3; i += BQ_THREADS) tile[i] = pb[(size_t)t0 * 3 + i];
        __syncthreads();
        if (__syncthreads_and(cnt >= nsample)) break;
        for (int j = 0; j < tn && cnt < nsample; ++j) {
            const float d2 = direct_dist(qx - tile[j * 3 + 0], qy - tile[j * 3 + 1], qz - tile[j * 3 + 2]);
            if (d2 < r2) {
                if (cnt == 0)
                    for (int l = 0; l < nsample; ++l) o[l] = t0 + j;
                o[cnt++] = t0 + j;
            }
        }
    }
    if (active && cnt == 0)
        for (int l = 0; l < nsample; ++l) o[l] = 0;       // pointnet2_utils.py:224 pre-zeroes idx
}

}  // namespace kdpc

using namespace kdpc;

KDPC_API int kdpc_gather(int b, int c, int n, int m, const float *f, const int *idx, float *out, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(f && idx && out && b > 0 && c > 0 && n > 0 && m > 0);
    if (b > 65535 || c > 65535) return KDPC_EUNSUPPORTED;
    dim3 grid((m + 255) / 256, c, b);
    gather_cm_kernel<<<grid, 256, 0, to_stream(stream)>>>(c, n, m, f, idx, out);
    KDPC_RETURN_LAST();
}

KDPC_API int kdpc_group(int b, int c, int n, int s, int k, const float *f, const int *idx, float *out,
                        kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(f && idx && out && b > 0 && c > 0 && n > 0 && s > 0 && k > 0);
    if (b > 65535) return KDPC_EUNSUPPORTED;
    const long long skl = (long long)s * k;
    if (skl > 0x7fffffffLL) return KDPC_EUNSUPPORTED;
    const int sk = (int)skl;
    cudaStream_t st = to_stream(stream);
    constexpr int CPB = 2;
    const size_t smem = (size_t)CPB * n * sizeof(float);
    if (smem <= 72 * 1024 && c >= 2 && skl >= 4096) {
        KDPC_ENSURE_SMEM((group_cm_smem_kernel<CPB>), 72 * 1024);
        const int cgroups = (c + CPB - 1) / CPB;
        // ~2 waves of 3 CTAs per SM; every CTA must amortise staging its channel rows over >= n gathered entries
        int split = (6 * device_sms() + cgroups * b - 1) / (cgroups * b);
        split = max(1, min(split, (int)(skl / (long long)n)));
        split = max(split, 1);
        int chunk = (sk + split - 1) / split;
        chunk = (chunk + 3) & ~3;
        split = (sk + chunk - 1) / chunk;
        dim3 grid(split, cgroups, b);
        group_cm_smem_kernel<CPB><<<grid, 256, smem, st>>>(c, n, sk, chunk, f, idx, out);
    } else {
        dim3 grid((sk + 255) / 256, b);
        group_cm_direct_kernel<<<grid, 256, 0, st>>>(c, n, sk, f, idx, out);
    }
    KDPC_RETURN_LAST();
}

KDPC_API int kdpc_gather_rows(int b, int n, int m, int c, const float *f, const int *idx, float *out,
                              kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(f && idx && out && b > 0 && n > 0 && m > 0 && c > 0);
    cudaStream_t st = to_stream(stream);
    const bool vec = (c % 4 == 0) && ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
    if (vec) {
        const long long total = (long long)b * m * (c / 4);
        if (total + 1024 < (1ll << 32)) {
            gather_rows4_kernel<<<(unsigned)div_up_ll(total, 1024), 256, 0, st>>>(
                (unsigned)total, (unsigned)n, (unsigned)m, (unsigned)(c / 4), reinterpret_cast<const float4 *>(f), idx,
                reinterpret_cast<float4 *>(out));
            KDPC_RETURN_LAST();
        }
        gather_rows_kernel<4><<<(unsigned)div_up_ll(total, 256), 256, 0, st>>>(total, n, m, c / 4, f, idx, out);
    } else {
        const long long total = (long long)b * m * c;
        gather_rows_kernel<1><<<(unsigned)div_up_ll(total, 256), 256, 0, st>>>(total, n, m, c, f, idx, out);
    }
    KDPC_RETURN_LAST();
}

// ---- channel concatenation of up to four row-strided matrices (torch.cat(..., dim = channels), pointconv_util.py:2242) ----
// One thread per 16-byte piece of the output; the sources may be column blocks of wider tensors (row strides).  torch's
// own cat falls back to a scalar kernel as soon as one input is a strided view (35 us for the 33 MB level-0 tensor).
struct ConcatSrc {
    const float4 *p[4];
    unsigned ld4[4];        // row stride in 16-byte units
    unsigned end4[4];       // exclusive prefix sums of the widths, in 16-byte units
};
__global__ void __launch_bounds__(256)
concat_rows_kernel(unsigned total, unsigned w4, unsigned ldo4, const ConcatSrc s, float4 *__restrict__ out) {
    const unsigned e = blockIdx.x * 256u + threadIdx.x;
    if (e >= total) return;
    const unsigned row = e / w4, c = e - row * w4;
    const int k = (c >= s.end4[0]) + (c >= s.end4[1]) + (c >= s.end4[2]);
    const unsigned c0 = k ? s.end4[k - 1] : 0u;
    const float4 v = ld_stream_f4(s.p[k] + (size_t)row * s.ld4[k] + (c - c0));
    out[(size_t)row * ldo4 + c] = v;
}

KDPC_API int kdpc_concat_rows(long long rows, int nsrc, const float *const *src, const int *ld, const int *width, float *out,
                              int ldo, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(src && ld && width && out && rows >= 0 && nsrc >= 1 && nsrc <= 4);
    ConcatSrc s;
    unsigned w4 = 0;
    for (int i = 0; i < 4; ++i) {
        if (i < nsrc) {
            KDPC_CHECK_ARGS(src[i] && width[i] > 0 && ld[i] >= width[i]);
            if ((width[i] & 3) || (ld[i] & 3) || (reinterpret_cast<uintptr_t>(src[i]) & 15)) return KDPC_EUNSUPPORTED;
            s.p[i] = reinterpret_cast<const float4 *>(src[i]);
            s.ld4[i] = (unsigned)ld[i] / 4u;
            w4 += (unsigned)width[i] / 4u;
        } else {
            s.p[i] = s.p[nsrc - 1];
            s.ld4[i] = 0;
        }
        s.end4[i] = i < nsrc ? w4 : 0xffffffffu;
    }
    if ((ldo & 3) || (unsigned)ldo / 4u < w4 || (reinterpret_cast<uintptr_t>(out) & 15)) return KDPC_EUNSUPPORTED;
    const long long total = rows * (long long)w4;
    if (total >= (1ll << 32) - 256) return KDPC_EUNSUPPORTED;
    if (total == 0) return 0;
    concat_rows_kernel<<<(unsigned)div_up_ll(total, 256), 256, 0, to_stream(stream)>>>(
        (unsigned)total, w4, (unsigned)ldo / 4u, s, reinterpret_cast<float4 *>(out));
    KDPC_RETURN_LAST();
}

static int kdpc_group_concat_direct = 1;     // (0: always the staged kernel; tests compare the two)
extern "C" __attribute__((visibility("default"))) void kdpc_group_concat_set_direct(int on) { kdpc_group_concat_direct = on; }

KDPC_API int kdpc_group_concat(int b, int n, int s, int k, int d, const float *cand_xyz, const float *query_xyz,
                               const float *feats, const int *idx, float *out, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(cand_xyz && query_xyz && idx && out && b > 0 && n > 0 && s > 0 && k > 0 && d >= 0);
    KDPC_CHECK_ARGS(d == 0 || feats != nullptr);
    if ((reinterpret_cast<uintptr_t>(out) % 16) != 0) return KDPC_EINVAL;
    if (d > 0 && (d % 4 == 0) && (reinterpret_cast<uintptr_t>(feats) % 16) != 0) return KDPC_EINVAL;
    const long long rows = (long long)b * s * k;
    if (d >= 32 && rows + 1024 < (1ll << 32) && kdpc_group_concat_direct) {
        group_concat_rows_kernel<<<(unsigned)div_up_ll(rows, 8 * GCR), 256, 0, to_stream(stream)>>>(
            (unsigned)rows, (unsigned)n, (unsigned)s, (unsigned)k, (unsigned)d, cand_xyz, query_xyz, feats, idx, out);
        KDPC_RETURN_LAST();
    }
    const size_t smem = (size_t)GC_ROWS * (3 + d) * sizeof(float);
    if (smem > 200 * 1024) return KDPC_EUNSUPPORTED;
    KDPC_ENSURE_SMEM(group_concat_kernel, 200 * 1024);
    group_concat_kernel<<<(unsigned)div_up_ll(rows, GC_ROWS), GC_THREADS, smem, to_stream(stream)>>>(
        rows, n, s, k, d, cand_xyz, query_xyz, feats, idx, out);
    KDPC_RETURN_LAST();
}

KDPC_API int kdpc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                             const float *xyz, int *idx, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(new_xyz && xyz && idx && b > 0 && n > 0 && m > 0 && nsample > 0);
    if (b > 65535) return KDPC_EUNSUPPORTED;
    dim3 grid((m + BQ_THREADS - 1) / BQ_THREADS, b);
    ball_query_kernel<<<grid, BQ_THREADS, 0, to_stream(stream)>>>(n, m, radius, nsample, new_xyz, xyz, idx);
    KDPC_RETURN_LAST();
}
