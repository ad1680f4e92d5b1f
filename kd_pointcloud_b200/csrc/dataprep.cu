// Input pipeline on the GPU: ProcessData and Augmentation of the reference's loaders (transforms/transforms.py:137-316).
//
// The reference runs, per sample and on the host, a rigid augmentation of both clouds (scale . rotation, shift, jitter;
// a second rotation + shift of cloud 2), the depth mask  pc1.z < T and pc2.z < T,  np.where (a stable compaction) and
// np.random.choice of num_points survivors (fancy-index gathers) -- numpy passes over up to 130 k points per sample,
// which becomes the wall once the model itself runs at > 1000 pairs/s.  Here a BATCH of padded raw clouds is
// processed by three launches:
//   1. dataprep_transform_kernel : the two affine maps (fp32, matrices prepared by the caller exactly as the reference
//                                  builds them), flow = pc2' - pc1', optional second jitter, depth mask
//   2. dataprep_compact_kernel   : one CTA per sample, stable compaction of the mask (ballot + prefix sums): the list
//                                  np.where returns, and its length
//   3. dataprep_select_kernel    : out[j] = survivors[sel[j]]  (pc1 and flow by sel1, pc2 by sel2)
// Randomness stays with the caller: ``sel`` are POSITIONS in the survivor list (what numpy draws: choice(indices, n,
// replace=False) == indices[permutation(len(indices))[:n]], replace=True == indices[randint(0, len, n)]), so for
// given draws the result is defined and equal to the reference's.  Everything is HBM-bound streaming.
#include "common.cuh"

namespace kdpc {

struct PrepAffine {            // one sample's maps; all row-vector conventions of the reference (p' = p . M + t)
    float m1[9];               // together: scale . rot^T                         (transforms.py:229-243)
    float t1[3];               // together shift                                  (:246-248)
    float m2[9];               // pc2 only: rot2^T                                (:262-270)
    float t2[3];               // pc2 shift                                       (:272-276)
};

__device__ __forceinline__ void affine3(const float *m, float x, float y, float z, float &ox, float &oy, float &oz) {
    // numpy's float32 [n,3] . [3,3]: three products summed left to right, no fused contraction
    ox = __fadd_rn(__fadd_rn(__fmul_rn(x, m[0]), __fmul_rn(y, m[3])), __fmul_rn(z, m[6]));
    oy = __fadd_rn(__fadd_rn(__fmul_rn(x, m[1]), __fmul_rn(y, m[4])), __fmul_rn(z, m[7]));
    oz = __fadd_rn(__fadd_rn(__fmul_rn(x, m[2]), __fmul_rn(y, m[5])), __fmul_rn(z, m[8]));
}

// grid (ceil(nmax / 256), B).  pc*_raw [B,nmax,stride] (stride >= 3: loaders keep extra columns), n_raw [B].
// work: [B,nmax,9] = pc1' | pc2' | flow ;  mask [B,nmax] (0/1).
__global__ void __launch_bounds__(256)
dataprep_transform_kernel(int nmax, int stride, float depth_threshold, int augment, const int *__restrict__ n_raw,
                          const float *__restrict__ pc1_raw, const float *__restrict__ pc2_raw,
                          const PrepAffine *__restrict__ aff, const float *__restrict__ jitter1,
                          const float *__restrict__ jitter2, float *__restrict__ work, unsigned char *__restrict__ mask) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nmax) return;
    const size_t r = (size_t)b * nmax + i;
    if (i >= n_raw[b]) { mask[r] = 0; return; }
    const float *a = pc1_raw + r * stride, *c = pc2_raw + r * stride;
    float x1 = a[0], y1 = a[1], z1 = a[2], x2 = c[0], y2 = c[1], z2 = c[2];
    float fx, fy, fz;
    if (augment) {
        const PrepAffine &A = aff[b];
        float bx = A.t1[0], by = A.t1[1], bz = A.t1[2];
        if (jitter1 != nullptr) {                                  // bias = shifts + jitter      (transforms.py:255)
            bx = __fadd_rn(bx, jitter1[r * 3 + 0]); by = __fadd_rn(by, jitter1[r * 3 + 1]); bz = __fadd_rn(bz, jitter1[r * 3 + 2]);
        }
        float tx, ty, tz;
        affine3(A.m1, x1, y1, z1, tx, ty, tz);
        x1 = __fadd_rn(tx, bx); y1 = __fadd_rn(ty, by); z1 = __fadd_rn(tz, bz);            // :257
        affine3(A.m1, x2, y2, z2, tx, ty, tz);
        x2 = __fadd_rn(tx, bx); y2 = __fadd_rn(ty, by); z2 = __fadd_rn(tz, bz);            // :258
        affine3(A.m2, x2, y2, z2, tx, ty, tz);
        x2 = __fadd_rn(tx, A.t2[0]); y2 = __fadd_rn(ty, A.t2[1]); z2 = __fadd_rn(tz, A.t2[2]);   // :278
        fx = __fsub_rn(x2, x1); fy = __fsub_rn(y2, y1); fz = __fsub_rn(z2, z1);            // :279 (before jitter2)
        if (jitter2 != nullptr) {                                  // :281-285 (only when the clouds correspond)
            x2 = __fadd_rn(x2, jitter2[r * 3 + 0]); y2 = __fadd_rn(y2, jitter2[r * 3 + 1]); z2 = __fadd_rn(z2, jitter2[r * 3 + 2]);
        }
    } else {
        fx = __fsub_rn(x2, x1); fy = __fsub_rn(y2, y1); fz = __fsub_rn(z2, z1);            // ProcessData :149
    }
    float *w = work + r * 9;
    w[0] = x1; w[1] = y1; w[2] = z1; w[3] = x2; w[4] = y2; w[5] = z2; w[6] = fx; w[7] = fy; w[8] = fz;
    mask[r] = (depth_threshold > 0.f) ? ((z1 < depth_threshold && z2 < depth_threshold) ? 1 : 0) : 1;   // :151-154, :287-290
}

// One CTA of 1024 threads per sample: survivors[b, 0..count) = ascending positions with mask == 1 (np.where), count[b].
__global__ void __launch_bounds__(1024)
dataprep_compact_kernel(int nmax, const unsigned char *__restrict__ mask, int *__restrict__ survivors, int *__restrict__ count) {
    __shared__ int warp_sums[32];
    __shared__ int base_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned char *mb = mask + (size_t)b * nmax;
    int *sb = survivors + (size_t)b * nmax;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int i0 = 0; i0 < nmax; i0 += 1024) {
        const int i = i0 + tid;
        const bool keep = i < nmax && mb[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_sums[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        if (warp == 0) {
            const int v = warp_sums[lane];
            int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            warp_sums[lane] = inc - v;                              // exclusive prefix of the warp totals
            total = __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncthreads();
        before = base_s + warp_sums[warp] + __popc(bal & ((1u << lane) - 1u));
        if (keep) sb[before] = i;
        __syncthreads();
        if (tid == 0) base_s += total;
        __syncthreads();
    }
    if (tid == 0) count[b] = base_s;
}

// grid (ceil(num_points / 256), B): out rows gathered through the caller's draws (positions in the survivor list).
// status[b] |= 1 when a draw is out of range (>= count[b]): the rows are then zero.
__global__ void __launch_bounds__(256)
dataprep_select_kernel(int nmax, int num_points, const float *__restrict__ work, const int *__restrict__ survivors,
                       const int *__restrict__ count, const int *__restrict__ sel1, const int *__restrict__ sel2,
                       float *__restrict__ out_pc1, float *__restrict__ out_pc2, float *__restrict__ out_sf,
                       int *__restrict__ status) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= num_points) return;
    const int c = count[b];
    const size_t o = ((size_t)b * num_points + j) * 3;
    const int s1 = sel1[(size_t)b * num_points + j], s2 = sel2[(size_t)b * num_points + j];
    if (s1 < 0 || s1 >= c || s2 < 0 || s2 >= c) {
        atomicOr(status + b, 1);
        for (int d = 0; d < 3; ++d) { out_pc1[o + d] = 0.f; out_pc2[o + d] = 0.f; out_sf[o + d] = 0.f; }
        return;
    }
    const float *w1 = work + ((size_t)b * nmax + survivors[(size_t)b * nmax + s1]) * 9;
    const float *w2 = work + ((size_t)b * nmax + survivors[(size_t)b * nmax + s2]) * 9;
    out_pc1[o + 0] = w1[0]; out_pc1[o + 1] = w1[1]; out_pc1[o + 2] = w1[2];     // pc1[sampled_indices1]   (:190, :313)
    out_sf[o + 0] = w1[6]; out_sf[o + 1] = w1[7]; out_sf[o + 2] = w1[8];        // sf[sampled_indices1]
    out_pc2[o + 0] = w2[3]; out_pc2[o + 1] = w2[4]; out_pc2[o + 2] = w2[5];     // pc2[sampled_indices2]
}

}  // namespace kdpc

using namespace kdpc;

/* bytes of the workspace of kdpc_dataprep_mask (transformed clouds + flow, mask, survivor list) */
KDPC_API long long kdpc_dataprep_workspace_bytes(int b, int nmax) {
    if (b <= 0 || nmax <= 0) return 0;
    const long long rows = (long long)b * nmax;
    return rows * 9 * 4 + ((rows + 15) / 16) * 16 + rows * 4;
}

/* Step 1 + 2 (transforms.py:149-156 / :229-293): transform (augment != 0), mask, compact.  affine: [B] records of 24 floats
 * (m1[9] t1[3] m2[9] t2[3]) or NULL when augment == 0; jitter1 / jitter2 [B,nmax,3] or NULL.  count [B] receives the
 * number of survivors per sample (len(np.where(mask)[0])). */
KDPC_API int kdpc_dataprep_mask(int b, int nmax, int stride, float depth_threshold, int augment, const int *n_raw,
                                const float *pc1_raw, const float *pc2_raw, const float *affine, const float *jitter1,
                                const float *jitter2, void *ws, int *count, kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(b > 0 && nmax > 0 && stride >= 3 && n_raw && pc1_raw && pc2_raw && ws && count);
    KDPC_CHECK_ARGS(!augment || affine != nullptr);
    if (b > 65535 || (reinterpret_cast<uintptr_t>(ws) % 16) != 0) return KDPC_EINVAL;
    const long long rows = (long long)b * nmax;
    float *work = reinterpret_cast<float *>(ws);
    unsigned char *mask = reinterpret_cast<unsigned char *>(ws) + rows * 36;
    int *survivors = reinterpret_cast<int *>(mask + ((rows + 15) / 16) * 16);
    cudaStream_t st = to_stream(stream);
    dim3 grid((nmax + 255) / 256, b);
    dataprep_transform_kernel<<<grid, 256, 0, st>>>(nmax, stride, depth_threshold, augment, n_raw, pc1_raw, pc2_raw,
                                                    reinterpret_cast<const PrepAffine *>(affine), jitter1, jitter2, work, mask);
    dataprep_compact_kernel<<<b, 1024, 0, st>>>(nmax, mask, survivors, count);
    KDPC_RETURN_LAST();
}

/* Step 3 (:158-192 / :295-315): gather num_points rows through the draws.  sel1 / sel2 [B,num_points]: positions in the
 * survivor list (sel2 == sel1 when the clouds correspond, NO_CORR false).  status [B] must be zeroed by the caller;
 * bit 0 is set for a sample with an out-of-range draw. */
KDPC_API int kdpc_dataprep_select(int b, int nmax, int num_points, const void *ws, const int *count, const int *sel1,
                                  const int *sel2, float *out_pc1, float *out_pc2, float *out_sf, int *status,
                                  kdpc_stream_t stream) {
    KDPC_CHECK_ARGS(b > 0 && nmax > 0 && num_points > 0 && ws && count && sel1 && sel2 && out_pc1 && out_pc2 && out_sf && status);
    if (b > 65535) return KDPC_EUNSUPPORTED;
    const long long rows = (long long)b * nmax;
    const float *work = reinterpret_cast<const float *>(ws);
    const unsigned char *mask = reinterpret_cast<const unsigned char *>(ws) + rows * 36;
    const int *survivors = reinterpret_cast<const int *>(mask + ((rows + 15) / 16) * 16);
    dim3 grid((num_points + 255) / 256, b);
    dataprep_select_kernel<<<grid, 256, 0, to_stream(stream)>>>(nmax, num_points, work, survivors, count, sel1, sel2,
                                                              out_pc1, out_pc2, out_sf, status);
    KDPC_RETURN_LAST();
}
