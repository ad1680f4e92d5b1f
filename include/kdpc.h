/*
 * kdpc.h — C ABI of libkdpc.so, the B200 (sm_100a) point-cloud op library.
 *
 * This is the drop-in boundary for the hot path of yunminjin2/KD-PointCloud:
 * every entry point replaces one native launcher of the reference's
 * `pointnet2_cuda` extension (pointnet2/src/*_gpu.h, bound in
 * pointnet2/src/pointnet2_api.cpp:10-24) or one torch op chain of
 * pointconv_util.py that the reference runs as many small library kernels.
 *
 * Conventions (same as the reference's launchers, sampling_gpu.h:12-27):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer;
 *   - the CALLER allocates every output and workspace (pointnet2_utils.py:25-26,
 *     55,94-95,128,172); kernels never allocate;
 *   - work is enqueued on `stream`; nothing synchronises;
 *   - return value: 0 on success, a positive cudaError_t on a launch failure,
 *     a negative KDPC_E* code for a rejected argument.  (The reference prints to
 *     stderr and calls exit(-1), sampling_gpu.cu:248-252; we return instead.)
 *   - "cm" tensors are channel-major  [B, C, N]  (the pointnet2 layout);
 *     "pm" tensors are point-major    [B, N, C]  (what pointconv_util.py permutes to
 *     before every op, e.g. pointconv_util.py:118,130).
 */
#ifndef KDPC_H_
#define KDPC_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *kdpc_stream_t;

#define KDPC_OK            0
#define KDPC_EINVAL       -1   /* null pointer / non-positive size                    */
#define KDPC_EUNSUPPORTED -2   /* size outside what the kernels are instantiated for  */

#define KDPC_ABI_VERSION 1
int kdpc_abi_version(void);
const char *kdpc_error_string(int code);
/* SM budget of the persistent (one CTA per SM) kernels and of the split plans: 0 = the whole device (default), n > 0 = at
 * most n SMs, so that a concurrent stream keeps the rest (the runner overlaps the next batch's sampling pyramid). */
void kdpc_set_sm_limit(int n);
int kdpc_sm_limit(void);

/* ---- input pipeline (transforms/transforms.py:137-316: ProcessData, Augmentation) ---------------------------------
 * A batch of padded raw clouds pc*_raw [B,nmax,stride>=3] with n_raw[B] valid points each.
 * kdpc_dataprep_mask: optional augmentation (affine[b] = m1[9] t1[3] m2[9] t2[3] as the reference builds them,
 *   jitter1 / jitter2 [B,nmax,3] or NULL), flow = pc2' - pc1', depth mask (pc1.z < T && pc2.z < T; T <= 0: all),
 *   stable compaction (np.where) into the workspace; count[b] = number of survivors.
 * kdpc_dataprep_select: out[b,j] = survivors[sel[b,j]] for caller-supplied draws sel1 / sel2 (positions in the
 *   survivor list, exactly what np.random.choice(indices, n, replace) draws); status[b] bit 0 = draw out of range. */
long long kdpc_dataprep_workspace_bytes(int b, int nmax);
int kdpc_dataprep_mask(int b, int nmax, int stride, float depth_threshold, int augment, const int *n_raw,
                       const float *pc1_raw, const float *pc2_raw, const float *affine, const float *jitter1,
                       const float *jitter2, void *ws, int *count, kdpc_stream_t stream);
int kdpc_dataprep_select(int b, int nmax, int num_points, const void *ws, const int *count, const int *sel1,
                         const int *sel2, float *out_pc1, float *out_pc2, float *out_sf, int *status,
                         kdpc_stream_t stream);

/* ---- pointnet2 ops (channel-major API of pointnet2/pointnet2_utils.py) ----------------- */

/* furthest_point_sampling_kernel_launcher, sampling_gpu.h:26-27 / sampling_gpu.cu:93-253.
 * xyz [B,N,3] -> idx int32 [B,M], idx[:,0] = 0.  temp [B,N] may be NULL; when given it
 * receives the final min-distance field exactly as the reference leaves it. Bit-exact
 * incl. the reference's block-size dependent tie rule. */
int kdpc_fps(int b, int n, int m, const float *xyz, float *temp, int *idx, kdpc_stream_t stream);
/* Clouds of 2048..16384 points run on a cluster of 8 CTAs per cloud (distributed-shared-memory arg-max);
 * kdpc_fps_set_cluster(0) forces the one-CTA-per-cloud kernel (same results; for A/B measurements). */
void kdpc_fps_set_cluster(int on);
/* Measurement hooks (tools/bench_fps.py): kdpc_fps_set_cluster(spread * 1000000 + ctas * 10000 + threads_per_cloud)
 * forces one cluster shape (ctas in {2,4,8}, threads_per_cloud in {1024,2048}; spread = 1 asks for
 * cudaClusterSchedulingPolicySpread); kdpc_fps_cluster_capacity = clusters of that shape that fit on the device at once. */
int kdpc_fps_cluster_capacity(int ctas, int threads_per_cloud, int n);

/* gather_points_kernel_launcher_fast, sampling_gpu.h:12-13.  f [B,C,N], idx [B,M] -> out [B,C,M] */
int kdpc_gather(int b, int c, int n, int m, const float *f, const int *idx, float *out, kdpc_stream_t stream);
/* gather_points_grad_kernel_launcher_fast, sampling_gpu.h:19-20. grad_f [B,C,N] is OVERWRITTEN
 * (no pre-zeroing needed) and the sum order is fixed (ascending j) => deterministic, unlike the
 * reference's atomicAdd.  ws: B*(N+1+M)*4 bytes (inverse index, see kdpc_build_csr). */
int kdpc_gather_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, void *ws,
                     float *grad_f, kdpc_stream_t stream);

/* group_points_kernel_launcher_fast, group_points_gpu.h.  f [B,C,N], idx [B,S,K] -> out [B,C,S,K] */
int kdpc_group(int b, int c, int n, int s, int k, const float *f, const int *idx, float *out, kdpc_stream_t stream);
/* group_points_grad_kernel_launcher_fast. grad_f [B,C,N] overwritten, deterministic.
 * ws: B*(N+1+S*K)*4 bytes. */
int kdpc_group_grad(int b, int c, int n, int s, int k, const float *grad_out, const int *idx, void *ws,
                    float *grad_f, kdpc_stream_t stream);

/* three_nn_kernel_launcher_fast, interpolate_gpu.h. unknown [B,N,3], known [B,M,3] ->
 * dist2 [B,N,3] (SQUARED, ascending), idx int32 [B,N,3]; ties keep the lower index.
 * ws: kdpc_knn_workspace_bytes(b, n, m) bytes. */
int kdpc_three_nn(int b, int n, int m, const float *unknown, const float *known, void *ws,
                  float *dist2, int *idx, kdpc_stream_t stream);
/* three_interpolate_kernel_launcher_fast. f [B,C,M], idx/w [B,N,3] -> out [B,C,N] */
int kdpc_three_interpolate(int b, int c, int m, int n, const float *f, const int *idx, const float *w,
                           float *out, kdpc_stream_t stream);
/* three_interpolate_grad_kernel_launcher_fast (n interpolated points, m source points, as in
 * interpolate_gpu.h). grad_f [B,C,M] overwritten, deterministic.  ws: B*(M+1+3N)*4 bytes. */
int kdpc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                const float *w, void *ws, float *grad_f, kdpc_stream_t stream);
/* ball_query_kernel_launcher_fast, ball_query_gpu.h. idx [B,M,nsample] fully written. */
int kdpc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                    const float *xyz, int *idx, kdpc_stream_t stream);

/* ---- pointconv_util ops (point-major) ------------------------------------------------ */

/* square_distance, pointconv_util.py:73-94: out[b,i,j] = rn(rn(-2*dot + |src_i|^2) + |dst_j|^2) */
int kdpc_square_distance(int b, int s, int n, const float *src, const float *dst, float *out, kdpc_stream_t stream);

/* kNN in a C-dimensional FEATURE space: square_distance + topk on [B,*,C] feature tensors, as CrossLayerLightFG uses them
 * (pointconv_util.py:1905 knn_point(nsample//2, knn2, knn1); square_distance :73-94, topk :106).
 * query [B,S,C], cand [B,N,C] -> idx int32 [B,S,k] ascending (distance, index); dist [B,S,k] or NULL.  k <= 32, C <= 512. */
int kdpc_knn_feat(int b, int s, int n, int c, int k, const float *query, const float *cand, int *idx, float *dist,
                  kdpc_stream_t stream);

/* knn_point, pointconv_util.py:96-107, without materialising the [B,S,N] matrix.
 * query [B,S,3], cand [B,N,3] -> the k candidates with smallest square_distance, ordered
 * ascending by (distance, index).  idx32 / idx64 / dist may each be NULL.  k <= min(32, N).
 * ws: kdpc_knn_workspace_bytes(b, s, n) bytes, 16-byte aligned.
 * For 256 <= N <= 16384 this runs kdpc_spatial_sort on both clouds (once if query == cand) and then
 * kdpc_knn_sorted; otherwise the brute-force kernel.  Results are identical either way. */
long long kdpc_knn_workspace_bytes(int b, int s, int n);
int kdpc_knn(int b, int s, int n, int k, const float *query, const float *cand, void *ws,
             int *idx32, long long *idx64, float *dist, kdpc_stream_t stream);
/* The brute-force kernel alone (every query scans every candidate through TMA-staged shared-memory
 * tiles).  ws: B*N*16 bytes. */
int kdpc_knn_bruteforce(int b, int s, int n, int k, const float *query, const float *cand, void *ws,
                        int *idx32, long long *idx64, float *dist, kdpc_stream_t stream);

/* Spatial (Morton) sort of B clouds of n <= 16384 points: out receives, per cloud, the points in
 * Morton order as float4 (x,y,z,|p|^2), one bounding box per tile of 64 consecutive points and the
 * original indices.  out: kdpc_spatial_sort_bytes(b, n) bytes, 16-byte aligned.  A sorted cloud can be
 * reused by any number of kdpc_knn_sorted calls (as queries or as candidates). */
long long kdpc_spatial_sort_bytes(int b, int n);
int kdpc_spatial_sort(int b, int n, const float *xyz, void *out, kdpc_stream_t stream);
/* The same representation for a DISPLACED copy of an already sorted cloud (a warped cloud, pointconv_util.py:2114-2142):
 * the parent's order is reused, the tile boxes are recomputed from the new coordinates; kdpc_knn_sorted results are
 * unchanged (the search needs valid boxes, not a good order). */
int kdpc_spatial_reorder(int b, int n, const float *xyz, const void *parent_ws, void *ws, kdpc_stream_t stream);
/* Exact kNN between two sorted clouds by best-first search over candidate tiles with a conservative
 * distance bound (knn_bf.cu).  direct = 0: square_distance rounding (knn_point); direct = 1: the
 * pointnet2 kernels' (dx^2+dy^2+dz^2) rounding (three_nn).  Output rows are in ORIGINAL query order. */
/* K <= 4 with >= 64 queries per SM: one thread per query, 32 Morton-consecutive queries share one walk over the candidate
 * tiles (default 1); 0 = one warp per query as for larger K.  Same results, bit for bit. */
void kdpc_knn_set_few(int on);
int kdpc_knn_sorted(int b, int s, int n, int k, int direct, const void *query_sorted, const void *cand_sorted,
                    int *idx32, long long *idx64, float *dist, kdpc_stream_t stream);

/* index_points_gather, pointconv_util.py:109-120 (pm): out[b,j,:] = f[b,idx[b,j],:]; f [B,N,C].
 * With m = S*K this is also index_points_group (pointconv_util.py:122-133) producing
 * [B,S,K,C] directly (the reference returns a permuted view of [B,C,S,K]). */
int kdpc_gather_rows(int b, int n, int m, int c, const float *f, const int *idx, float *out, kdpc_stream_t stream);

/* torch.cat((a, b, ...), dim = channels) of point-major activations - SceneFlowEstimatorResidual.forward,
 * pointconv_util.py:2242 (`torch.cat([feats, cost_volume], dim = 1)`) and the feature concatenations of
 * models_bid_pointconv.py:150-190: out[r, :] = [src[0][r, :width[0]] | src[1][r, :width[1]] | ...] for r < rows.
 * src / ld / width are HOST arrays of nsrc <= 4 entries (device pointers, row strides and widths in floats); a source may
 * be a column block of a wider tensor.  Widths, strides % 4 == 0 and 16-byte aligned pointers, else KDPC_EUNSUPPORTED. */
int kdpc_concat_rows(long long rows, int nsrc, const float *const *src, const int *ld, const int *width, float *out,
                     int ldo, kdpc_stream_t stream);

/* group / group_query, pointconv_util.py:135-182, fused: out[b,s,k,:] =
 * [cand_xyz[idx]-query_xyz[s] (3), feats[idx] (D)].  feats may be NULL (d = 0). */
/* A/B switch for measurements and tests: 0 = always the shared-memory staged kernel (same results). */
void kdpc_group_concat_set_direct(int on);
int kdpc_group_concat(int b, int n, int s, int k, int d, const float *cand_xyz, const float *query_xyz,
                      const float *feats, const int *idx, float *out, kdpc_stream_t stream);

/* WeightNet, pointconv_util.py:184-215 with hidden_unit=[h1,h2], bn=False: per row
 * relu(W3 relu(W2 relu(W1 x + b1) + b2) + b3).  x = in + row*in_stride (3 floats);
 * out [rows, wout].  Weights are the nn.Conv2d 1x1 weights flattened row-major [out,in]. */
int kdpc_weightnet(long long rows, const float *in, int in_stride, int h1, int h2, int wout,
                   const float *w1, const float *b1, const float *w2, const float *b2,
                   const float *w3, const float *b3, float *out, kdpc_stream_t stream);

/* Backward of WeightNet(3 -> 8 -> 8 -> wout), wout in {8, 16} (pointconv_util.py:184-215 under autograd): one kernel
 * recomputes the forward per row, back-propagates g_out [rows,wout] and reduces the six parameter gradients
 * deterministically (per-warp partials in ws, kdpc_weightnet_grad_ws_bytes).  g_in [rows,3] (optional): gradient w.r.t.
 * the localized coordinates. */
long long kdpc_weightnet_grad_ws_bytes(long long rows);
int kdpc_weightnet_grad(long long rows, const float *in, int in_stride, int wout, const float *g_out,
                        const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                        const float *b3, void *ws, float *gw1, float *gb1, float *gw2, float *gb2, float *gw3,
                        float *gb3, float *g_in, kdpc_stream_t stream);

/* PointConv aggregation, pointconv_util.py:249 / :437: out[r, c*wout + w] =
 * sum_k grouped[r,k,c] * wn[r,k,w].  grouped [R,K,C], wn [R,K,wout] -> out [R, C*wout]. */
int kdpc_pointconv_agg(long long rows, int k, int c, int wout, const float *grouped, const float *wn,
                       float *out, kdpc_stream_t stream);
/* Its backward (autograd of pointconv_util.py:249; torch runs two bmm of B*S tiny matrices): grad_out [R, C*wout] ->
 * grad_grouped[r,k,c] = sum_w wn[r,k,w] grad_out[r,c,w] and grad_wn[r,k,w] = sum_c grouped[r,k,c] grad_out[r,c,w]
 * (either may be NULL).  wout = 16, k <= 16.  Deterministic. */
int kdpc_pointconv_agg_grad(long long rows, int k, int c, int wout, const float *grouped, const float *wn,
                            const float *grad_out, float *grad_grouped, float *grad_wn, kdpc_stream_t stream);

/* CrossLayerLight.cross front half, pointconv_util.py:1836-1843:
 * out[b,s,k,:] = act(p2[b,idx[b,s,k],:] + p1[b,s,:] + pos_w (xyz2[idx]-xyz1[s]) + pos_b),
 * act = LeakyReLU(slope) (slope = 0 gives ReLU).  p1 [B,S,D], p2 [B,N,D], pos_w [D,3]. */
int kdpc_costvol_pre(int b, int s, int n, int k, int d, const float *xyz1, const float *xyz2,
                     const float *p1, const float *p2, const int *idx, const float *pos_w,
                     const float *pos_b, float slope, float *out, kdpc_stream_t stream);

/* max over the K axis: in [R,K,D] -> out [R,D], argmax int32 [R,D] (may be NULL). F.max_pool2d at
 * pointconv_util.py:1848. */
int kdpc_max_over_k(long long rows, int k, int d, const float *in, float *out, int *arg, kdpc_stream_t stream);

/* UpsampleFlow / PointWarping inverse-distance interpolation, pointconv_util.py:2131-2139,
 * 2164-2171: w_j = (1/max(|cand[idx_j]-q|,1e-10)) / sum_j(...);  out[b,i,:] = sum_j w_j feat[b,idx_j,:].
 * q_xyz [B,N,3], c_xyz [B,S,3], idx [B,N,3], feat [B,S,C] -> out [B,N,C]; w_out [B,N,3] may be NULL. */
int kdpc_interp3(int b, int n, int s, int c, const float *q_xyz, const float *c_xyz, const int *idx,
                 const float *feat, float *out, float *w_out, kdpc_stream_t stream);

/* ---- tensor-core (tcgen05) layers ------------------------------------------------------- */

/* Pack fp32 weights [N, K_src] (nn.Linear / 1x1-conv layout) into the bf16 hi/lo SWIZZLE_128B chunk
 * images the tcgen05 kernels stream with bulk TMA.  mode 0: plain (K_packed = K_src).  mode 1:
 * PointConv.linear (pointconv_util.py:223,250) for the fused kernel: channel order [features(d),
 * dx,dy,dz,0] x wn, K_src = (d+3)*wn, K_packed = (d+4)*wn.  out: kdpc_packed_weight_bytes(N, K_packed). N <= 256. */
long long kdpc_packed_weight_bytes(int n, int k_packed);
/* The fused tcgen05 layers stage their gathers asynchronously (bulk copies two pipeline iterations ahead);
 * kdpc_tc_set_async(0) selects the synchronous register-staged producers (same results; A/B measurements). */
void kdpc_tc_set_async(int on);
/* 1 = tcgen05 kernels are launched with programmatic dependent launch (default 0: measured neutral): their CTAs set up while the previous
 * kernel drains and wait (griddepcontrol.wait) before touching global memory; CUDA-graph capture records programmatic
 * edges.  Same results. */
void kdpc_tc_set_pdl(int on);
int kdpc_tc_pdl_enabled(void);
/* Small-M layers (<= half the SMs in 128-row tiles) with >= 128 outputs: work items are column blocks over the whole K
 * (default 1: final results from the epilogue, no workspace); 0 = split-K + reduce as for the narrow layers. */
void kdpc_linear_set_split_n(int on);
/* debug (tools/trace_pointconv.py, tools/trace_costvol.py): while a device buffer of 200 x 16 int64 is set, CTA 0 of every
 * tcgen05 kernel writes per-iteration clock64 stamps of its producer, MMA and epilogue warps into it; NULL = off */
void kdpc_tc_set_trace(void *device_buffer);
void *kdpc_tc_trace_buffer(void);
int kdpc_tc_async_enabled(void);
int kdpc_pack_weight(int n, int k_src, int mode, int d, int wn, const float *w, void *out, kdpc_stream_t stream);

/* y[M, ldo] = clamp(leaky(x[M, ldx(K)] W^T * scale + shift, slope), lo, hi) + residual    (N <= 256).
 * scale / shift / residual may be NULL; slope = 1 disables the activation; lo > hi disables the clamp.
 * Replaces nn.Linear / 1x1 Conv1d / Conv2d (+ eval BatchNorm + LeakyReLU) of pointconv_util.py:20-54,250-256,
 * 1797-1821, 2229-2255.  x, out and wpacked must be 16-byte aligned.
 * ws: kdpc_linear_tc_ws_bytes(m,n,k) bytes or NULL.  With a workspace, layers with few 128-row tiles and a long K
 * (PointConv linears of the coarse levels: 1024 x 8240 -> 256) are split along K over the idle SMs and reduced
 * in a fixed order (deterministic). */
long long kdpc_linear_tc_ws_bytes(long long m, int n, int k);
int kdpc_linear_tc(long long m, int n, int k, const float *x, int ldx, const void *wpacked,
                   const float *scale, const float *shift, float slope, float clamp_lo, float clamp_hi,
                   const float *residual, void *ws, float *out, int ldo, kdpc_stream_t stream);
/* Same contract on CUDA cores, for layers too small for a 128-row MMA tile (K < 16 or N < 16). w fp32 [N,K]. */
int kdpc_linear_simt(long long m, int n, int k, const float *x, int ldx, const float *w,
                     const float *scale, const float *shift, float slope, float clamp_lo, float clamp_hi,
                     const float *residual, float *out, int ldo, kdpc_stream_t stream);

/* Weight gradient of y = x W^T (nn.Linear / 1x1 conv; reference: autograd of pointconv_util.py:20-54, 223, 250):
 * dw[n,k] = sum_m dy[m,n] * x[m,k] on tcgen05 with MN-major operand tiles (bf16 hi/lo, fp32 accumulation), the row range
 * split over CTAs and reduced in split order (deterministic).  ws: kdpc_linear_dw_ws_bytes(m, n, k) bytes. */
/* narrow contiguous operands (n <= 128, k + 1 <= 128, ldy == n, ldx == k, m % 64 == 0): row chunks staged by bulk TMA
 * several chunks ahead, same bits as the register-staged fetch; k + 1 <= 4 (3 -> D layers): weighted column sums on the
 * CUDA cores in fp32 (no bf16 split: last-bit differences).  Default 1; 0 = the register-staged tcgen05 path for every shape. */
void kdpc_linear_dw_set_async(int on);
long long kdpc_linear_dw_ws_bytes(long long m, int n, int k);
int kdpc_linear_dw(long long m, int n, int k, const float *dy, int ldy, const float *x, int ldx, void *ws,
                   float *dw, int lddw, float *db /* [N] bias gradient = column sums of dy, or NULL */, kdpc_stream_t stream);

/* PointConv (pointconv_util.py:231-258) fused end to end for inference: neighbour gather + relative xyz +
 * WeightNet(3->8->8->16, ReLU) + sum over K + Linear(16(d+3) -> n_out) + scale/shift (bias, eval BatchNorm) +
 * LeakyReLU(slope).  cand_xyz [B,N,3], query_xyz [B,S,3], feats [B,N,d] (d % 4 == 0), idx int32 [B,S,k]
 * (k = 9 or 16) -> out [B,S,n_out].  wn_params: HOST array of 248 floats (w1[8x3] b1[8] w2[8x8] b2[8]
 * w3[16x8] b3[16], nn.Conv2d layouts) passed to the kernel as launch parameters.  wpacked: the Linear
 * weight packed with kdpc_pack_weight(mode 1, d, 16).  Neither [B,S,k,3+d] nor [B,S,16(d+3)] touches HBM.
 * ws: kdpc_pointconv_fused_ws_bytes(b,s,k,d,n_out) bytes or NULL (split-K partial sums, as for kdpc_linear_tc, followed by
 * the WeightNet outputs [b*s, k, 16] of a small pre-pass; with NULL both happen inside the fused kernel, slower). */
long long kdpc_pointconv_fused_ws_bytes(int b, int s, int k, int d, int n_out);
/* operand pipeline stages of the fused PointConv (default 2: the rest of the SM's L1/shared array serves the
 * neighbour gathers; measured 5% faster end to end than 3). */
void kdpc_pointconv_set_stages(int n);
/* WeightNet pre-pass (a small kernel that leaves the WeightNet outputs in the workspace; needs ws != NULL): 1 (default) =
 * where it pays (>= 32768 rows, or split-K), 0 = never (evaluated inside the fused kernel), 2 = always.  Same results. */
void kdpc_pointconv_set_precompute(int on);
int kdpc_pointconv_fused(int b, int n, int s, int k, int d, int n_out, const float *cand_xyz,
                         const float *query_xyz, const float *feats, const int *idx, const float *wn_params,
                         const void *wpacked, const float *scale, const float *shift, float slope,
                         void *ws, float *out, kdpc_stream_t stream);
/* Same, processing the queries of every cloud in a caller-given order (same results, bit for bit): tile position i of
 * cloud b handles query row_order[b * order_stride + i] (a permutation of 0..s-1 per cloud, order_stride >= s).  With a
 * spatially coherent order - the Morton order kdpc_spatial_sort leaves in its workspace - the 128 queries of a tile
 * share most of their neighbours, so the neighbour gathers hit in L1 instead of L2.  Ignored for split-K shapes. */
int kdpc_pointconv_fused_ordered(int b, int n, int s, int k, int d, int n_out, const float *cand_xyz,
                                 const float *query_xyz, const float *feats, const int *idx, const float *wn_params,
                                 const void *wpacked, const float *scale, const float *shift, float slope,
                                 const int *row_order, int order_stride, void *ws, float *out, kdpc_stream_t stream);
/* int32 element offset of the sorted-position -> original-index table inside one cloud's block of a
 * kdpc_spatial_sort workspace, and the size of that block in int32 elements (= the order_stride to pass above) */
int kdpc_spatial_sort_order_offset(int n);
int kdpc_spatial_sort_order_stride(int n);

/* CrossLayerLight.cross (pointconv_util.py:1826-1850) with a single-layer mlp, fused: out[b,i,:] =
 * max_k leaky(W act(p2[idx[b,i,k]] + p1[b,i] + pos_w (xyz2[idx]-xyz1[i]) + pos_b) + bias, slope_post).
 * k must be 32, d % 8 == 0, d, d_out <= 256.  wpacked: kdpc_pack_weight(mode 0) of W [d_out, d].
 * xyz1 [B,S,3], xyz2 [B,N,3], p1 [B,S,d], p2 [B,N,d], idx int32 [B,S,32] -> out [B,S,d_out].
 * ws: kdpc_costvol_fused_ws_bytes(b,s,n,d) bytes (16-byte aligned) for the per-point features with the positional
 * encoding folded in; NULL selects the variant that evaluates the encoding per neighbour (no workspace). */
/* One Adam step over `count` fp32 tensors in one launch per 96 tensors (reference: torch.optim.Adam(lr, betas, eps,
 * weight_decay) + optimizer.step(), distilTrain.py:134-135, 182; torch's capturable arithmetic: L2 weight decay, bias
 * correction, no amsgrad).  params / grads / exp_avg / exp_avg_sq: HOST arrays of `count` device pointers (passed to the
 * kernel by value: a CUDA graph records them); sizes: HOST array of element counts; lr, step: DEVICE fp32 scalars - the
 * update uses t = *step + 1 and *step is incremented afterwards. */
int kdpc_adam_step(int count, const void *const *params, const void *const *grads, const void *const *exp_avg,
                   const void *const *exp_avg_sq, const long long *sizes, const float *lr, float beta1, float beta2,
                   float eps, float weight_decay, float *step, kdpc_stream_t stream);
/* Backward of kdpc_costvol_fused in its folded form (reference: autograd through CrossLayerLight.cross, pointconv_util.py:1826-1850,
 * by loss.backward(), distilTrain.py:180), k = 32, d = d_out = 32 or 64:  out[i,c] = act2(max_k (W act1(p2q[idx[i,k]] + p1q[i]) + bias)[c])
 * with p1q = points1 + pos_b - pos_w xyz1 and p2q = points2 + pos_w xyz2 (the caller folds and un-folds the positional layer).
 * One warp per point recomputes the row block and lets the gradient of out[i,c] through to the ONE neighbour that attains the
 * maximum (the first one, like torch.max).  grad_p1q [b,s,d]; grad_rows [b*s*k, d] = gradient of every gathered p2q row, to be
 * scattered with kdpc_scatter_rows_csr (deterministic; no float atomics); grad_w [d_out,d]; grad_b [d_out] or NULL.  w: fp32
 * [d_out,d] (not packed).  ws: kdpc_costvol_grad_ws_bytes() bytes. */
long long kdpc_costvol_grad_ws_bytes(void);
int kdpc_costvol_grad(int b, int s, int n, int k, int d, int d_out, const float *p1q, const float *p2q, const int *idx,
                      const float *w, const float *bias, float slope_pre, float slope_post, const float *grad_out,
                      void *ws, float *grad_p1q, float *grad_rows, float *grad_w, float *grad_b, kdpc_stream_t stream);
/* Cost volume at the 8192-point level (d = 32, 16 < d_out <= 32): two 128-row tiles per pipeline iteration against the
 * block-diagonal weight diag(W, W) (default 1); 0 = one tile per iteration at every level.  Same results. */
void kdpc_costvol_set_pairing(int on);
long long kdpc_costvol_fused_ws_bytes(int b, int s, int n, int d);
int kdpc_costvol_fused(int b, int s, int n, int k, int d, int d_out, const float *xyz1, const float *xyz2,
                       const float *p1, const float *p2, const int *idx, const float *pos_w,
                       const float *pos_b, float slope_pre, const void *wpacked, const float *bias,
                       float slope_post, void *ws, float *out, kdpc_stream_t stream);

/* ---- losses (loss_functions.py) ------------------------------------------------------ */

/* multiScaleLoss (loss_functions.py:6-25) against up to two targets at once, forward + gradient:
 *   *loss += sum_t weight[t] * sum_s alpha[s] * sum_{b,p} || pred_s[b,:,p] - target_t[b, chain_s(p), :] ||_2
 * where chain_s composes the FPS index lists fps_idx[s-1] .. fps_idx[0] (the reference's chained
 * index_points_gather of the ground truth, never materialised here).  weight[t] carries the 1/B of the
 * reference's mean over the batch (and gamma / beta of the distillation losses).
 * n, pred, grad_pred, fps_idx, alpha, target, weight are HOST arrays (of device pointers where applicable);
 * pred[s] / grad_pred[s]: [B,3,n[s]] (point_major = 0) or [B,n[s],3] (1); grad_pred may be NULL, otherwise
 * grad_pred[s] is OVERWRITTEN with d loss / d pred_s.  fps_idx[s]: int32 [B,n[s+1]], s < nscales-1.
 * target[t]: [B,n[0],3].  nscales <= 4, ntargets <= 2.  ws: kdpc_loss_workspace_bytes() bytes, zeroed ONCE
 * before its first use.  Deterministic (no floating-point atomics). */
long long kdpc_loss_workspace_bytes(void);
int kdpc_flow_loss(int b, int nscales, int point_major, const int *n, const float *const *pred,
                   float *const *grad_pred, const int *const *fps_idx, const float *alpha, int ntargets,
                   const float *const *target, const float *weight, void *ws, float *loss, kdpc_stream_t stream);
/* distillation hint term (loss_functions.py:92-93, 213-216): *loss += 0.5 * weight * sum (fs - ft)^2;
 * grad_fs (may be NULL) = weight * (fs - ft).  fs, ft, grad_fs: n floats in the same order. */
int kdpc_hint_loss(long long n, const float *fs, const float *ft, float weight, float *grad_fs, void *ws,
                   float *loss, kdpc_stream_t stream);

/* ---- evaluation metrics ------------------------------------------------ */

/* evaluate_3d + evaluate_2d (evaluation_utils.py:17-50) over one batch, with the 2-D flows of
 * geometry.get_batch_2d_flow(pc1, pc1+gt, pc1+pred) / project_3d_to_2d (utils/geometry.py:6-65) computed on the fly.
 * pred: [B,3,N] (point_major = 0, the model's output layout) or [B,N,3] (1); gt, pc1: [B,N,3]; calib: NULL
 * (FlyingThings3D intrinsics f=-1050, cx=479.5, cy=269.5) or [B,6] = f, cx, cy, constx, consty, constz per sample
 * (KITTI P_rect_02, geometry.py:25-38); pc1 may be NULL (no 2-D metrics).  out: 6 floats on the device = EPE3D,
 * Acc3DS, Acc3DR, Outliers3D, EPE2D, Acc2D (batch means, as the reference's per-batch AverageMeter updates).
 * Per-point arithmetic is the reference's float32 sequence with IEEE rounding, so the accuracy counts are exact;
 * deterministic.  ws: kdpc_flow_metrics_workspace_bytes() bytes, 8-byte aligned, zeroed ONCE before first use. */
long long kdpc_flow_metrics_workspace_bytes(void);
int kdpc_flow_metrics(int b, int n, int point_major, const float *pred, const float *gt, const float *pc1,
                      const float *calib, void *ws, float *out, kdpc_stream_t stream);

/* ---- deterministic backward plumbing ------------------------------------------------- */

/* Inverse of an index list: idx int32 [B,M] with values in [0,N)  ->  offsets int32 [B,N+1] and
 * perm int32 [B,M] such that perm[b, offsets[b,i] .. offsets[b,i+1]) lists, ASCENDING, the j with
 * idx[b,j] == i.  Replaces the float atomicAdd of sampling_gpu.cu:62, group_points_gpu.cu:24,
 * interpolate_gpu.cu:139-141 by a fixed-order segmented reduction. */
int kdpc_build_csr(int b, int n, int m, const int *idx, int *offsets, int *perm, kdpc_stream_t stream);

/* grad_f[b,i,:] (=|+=) sum_{p in seg(i)} wgt[b,perm[p]] * g[b, perm[p]/gdiv, :]   (point-major,
 * g [B, M/gdiv, C], wgt [B,M] or NULL).  gdiv = 1: backward of gather_rows / grouping;
 * gdiv = 3 with wgt: backward of interp3 w.r.t. feat. */
int kdpc_scatter_rows_csr(int b, int n, int m, int c, int gdiv, const float *g, const float *wgt,
                          const int *offsets, const int *perm, float *grad_f, int accumulate,
                          kdpc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KDPC_H_ */
