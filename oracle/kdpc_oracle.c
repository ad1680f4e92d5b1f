/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the reference's
 * point-cloud ops.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (kd_pointcloud_b200/) never does.
 *
 * Each function restates one reference kernel in plain scalar C with the SAME
 * floating-point expression order the reference's kernels have after nvcc -O2
 * contraction (checked in the SASS of the reference compiled for sm_100a:
 * FMUL dy,dy ; FFMA dx,dx,. ; FFMA dz,dz,.).  Build with -ffp-contract=off so
 * that gcc performs exactly the roundings written here.
 *
 * Pinned against: oracle/_ref (the reference's own CUDA launchers, compiled
 * from /root/reference/pointnet2/src where they lie) in tests/test_ref_kernels_gpu.py,
 * and against the reference's torch layers through tests/golden/ (see
 * tests/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* pointnet2/src/cuda_utils.h:9-14  opt_n_threads: largest power of two <= n, capped at 1024 */
static int ref_block_size(int n) {
    int p = 1;
    while ((p << 1) <= n && (p << 1) <= 1024) p <<= 1;
    return p;
}

/* squared distance as the reference kernels compute it (sampling_gpu.cu:133,
 * interpolate_gpu.cu:37): (x..)^2 + (y..)^2 + (z..)^2 contracted by nvcc to
 * fma(dz,dz, fma(dx,dx, rn(dy*dy))). */
static inline float ref_sqdist3(float dx, float dy, float dz) {
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* ------------------------------------------------------------------------- */
/* furthest point sampling: pointnet2/src/sampling_gpu.cu:86-209             */
/* literal simulation of the block: strided per-thread scan + left-biased    */
/* shared-memory tree ( __update keeps idx1 on ties )                        */
void oracle_fps(int b, int n, int m, const float *xyz, float *temp, int *idx) {
    if (m <= 0) return;
    const int bs = ref_block_size(n);
    float *dists = (float *)malloc(sizeof(float) * bs);
    int *dists_i = (int *)malloc(sizeof(int) * bs);
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * n * 3;
        float *t = temp + (size_t)bi * n;
        int *out = idx + (size_t)bi * m;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int tid = 0; tid < bs; ++tid) {
                int besti = 0;
                float best = -1.f;
                for (int k = tid; k < n; k += bs) {
                    float d = ref_sqdist3(p[k * 3 + 0] - x1, p[k * 3 + 1] - y1, p[k * 3 + 2] - z1);
                    float d2 = fminf(d, t[k]);
                    t[k] = d2;
                    if (d2 > best) { best = d2; besti = k; }
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = bs >> 1; s >= 1; s >>= 1) {
                for (int tid = 0; tid < s; ++tid) {
                    float v1 = dists[tid], v2 = dists[tid + s];
                    int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = v1 > v2 ? v1 : v2;      /* max(v1, v2) */
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
    }
    free(dists);
    free(dists_i);
}

/* closed form of the same tie rule: among maximal d2 choose the
 * smallest (bitreverse(k mod bs), k / bs).  Used to cross-check the literal
 * simulation above. */
void oracle_fps_closed_form(int b, int n, int m, const float *xyz, float *temp, int *idx) {
    if (m <= 0) return;
    const int bs = ref_block_size(n);
    int lg = 0;
    while ((1 << lg) < bs) ++lg;
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * n * 3;
        float *t = temp + (size_t)bi * n;
        int *out = idx + (size_t)bi * m;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            float best = -1.f;
            uint32_t bestkey = 0xffffffffu;
            int besti = 0;
            for (int k = 0; k < n; ++k) {
                float d = ref_sqdist3(p[k * 3 + 0] - x1, p[k * 3 + 1] - y1, p[k * 3 + 2] - z1);
                float d2 = fminf(d, t[k]);
                t[k] = d2;
                uint32_t tid = (uint32_t)k & (uint32_t)(bs - 1), r = 0;
                for (int q = 0; q < lg; ++q) r |= ((tid >> q) & 1u) << (lg - 1 - q);
                uint32_t key = (r << 16) | (uint32_t)(k >> lg);
                if (d2 > best || (d2 == best && key < bestkey)) { best = d2; bestkey = key; besti = k; }
            }
            old = besti;
            out[j] = old;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* gather: sampling_gpu.cu:8-24, grad :46-63                                 */
void oracle_gather(int b, int c, int n, int m, const float *f, const int *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int j = 0; j < m; ++j)
                out[((size_t)bi * c + ci) * m + j] = f[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + j]];
}
void oracle_gather_grad(int b, int c, int n, int m, const float *g, const int *idx, float *gf) {
    /* gf must be zero-filled by the caller (pointnet2_utils.py:67); sequential add order j ascending */
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int j = 0; j < m; ++j)
                gf[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + j]] += g[((size_t)bi * c + ci) * m + j];
}

/* group: group_points_gpu.cu:47-66, grad :8-25                              */
void oracle_group(int b, int c, int n, int np, int ns, const float *f, const int *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int s = 0; s < np; ++s)
                for (int k = 0; k < ns; ++k)
                    out[(((size_t)bi * c + ci) * np + s) * ns + k] =
                        f[((size_t)bi * c + ci) * n + idx[((size_t)bi * np + s) * ns + k]];
}
void oracle_group_grad(int b, int c, int n, int np, int ns, const float *g, const int *idx, float *gf) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int s = 0; s < np; ++s)
                for (int k = 0; k < ns; ++k)
                    gf[((size_t)bi * c + ci) * n + idx[((size_t)bi * np + s) * ns + k]] +=
                        g[(((size_t)bi * c + ci) * np + s) * ns + k];
}

/* ------------------------------------------------------------------------- */
/* three_nn: interpolate_gpu.cu:9-52 (double compares of float distances,    */
/* strict '<' cascade, init 1e40 / index 0).  Writes SQUARED distances; the  */
/* Python wrapper takes sqrt (pointnet2_utils.py:98).                        */
void oracle_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) {
    for (int bi = 0; bi < b; ++bi)
        for (int i = 0; i < n; ++i) {
            const float *u = unknown + ((size_t)bi * n + i) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            double b1 = 1e40, b2 = 1e40, b3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                float d = ref_sqdist3(u[0] - kn[k * 3 + 0], u[1] - kn[k * 3 + 1], u[2] - kn[k * 3 + 2]);
                if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
                else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
                else if (d < b3) { b3 = d; i3 = k; }
            }
            float *od = dist2 + ((size_t)bi * n + i) * 3;
            int *oi = idx + ((size_t)bi * n + i) * 3;
            od[0] = (float)b1; od[1] = (float)b2; od[2] = (float)b3;
            oi[0] = i1; oi[1] = i2; oi[2] = i3;
        }
}

/* three_interpolate: interpolate_gpu.cu:77-97; w0*p0 + w1*p1 + w2*p2 is
 * contracted by nvcc to fma(w2,p2, fma(w0,p0, rn(w1*p1))) (SASS checked). */
void oracle_three_interpolate(int b, int c, int m, int n, const float *f, const int *idx, const float *w, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *fp = f + ((size_t)bi * c + ci) * m;
            for (int i = 0; i < n; ++i) {
                const int *ii = idx + ((size_t)bi * n + i) * 3;
                const float *ww = w + ((size_t)bi * n + i) * 3;
                float t = ww[1] * fp[ii[1]];
                t = fmaf(ww[0], fp[ii[0]], t);
                out[((size_t)bi * c + ci) * n + i] = fmaf(ww[2], fp[ii[2]], t);
            }
        }
}
void oracle_three_interpolate_grad(int b, int c, int n, int m, const float *g, const int *idx, const float *w, float *gf) {
    /* interpolate_gpu.cu:120-142; gf zero-filled by caller */
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int i = 0; i < n; ++i) {
                const int *ii = idx + ((size_t)bi * n + i) * 3;
                const float *ww = w + ((size_t)bi * n + i) * 3;
                float go = g[((size_t)bi * c + ci) * n + i];
                float *dst = gf + ((size_t)bi * c + ci) * m;
                dst[ii[0]] += go * ww[0];
                dst[ii[1]] += go * ww[1];
                dst[ii[2]] += go * ww[2];
            }
}

/* ball_query: ball_query_gpu.cu:9-45 (idx pre-zeroed by pointnet2_utils.py:224) */
void oracle_ball_query(int b, int n, int m, float radius, int ns, const float *new_xyz, const float *xyz, int *idx) {
    const float r2 = radius * radius;
    for (int bi = 0; bi < b; ++bi)
        for (int i = 0; i < m; ++i) {
            const float *q = new_xyz + ((size_t)bi * m + i) * 3;
            const float *p = xyz + (size_t)bi * n * 3;
            int *o = idx + ((size_t)bi * m + i) * ns;
            int cnt = 0;
            for (int k = 0; k < n && cnt < ns; ++k) {
                /* (a*a + b*b) + c*c contracted as in the other kernels:
                 * SASS of ball_query shows FMUL (second term) ; FFMA (first) ; FFMA (third) */
                float d2 = ref_sqdist3(q[0] - p[k * 3 + 0], q[1] - p[k * 3 + 1], q[2] - p[k * 3 + 2]);
                if (d2 < r2) {
                    if (cnt == 0) for (int l = 0; l < ns; ++l) o[l] = k;
                    o[cnt++] = k;
                }
            }
        }
}

/* ------------------------------------------------------------------------- */
/* square_distance: pointconv_util.py:73-94 as torch evaluates it in fp32:    */
/*   dot  = fma(z,z', fma(y,y', rn(x*x')))         (sgemm with k = 3)          */
/*   qq   = rn(rn(rn(x*x)+rn(y*y)) + rn(z*z))      (src**2 then sum(-1))       */
/*   dist = rn(rn(-2*dot + qq) + cc)               (two in-place adds)         */
/* Verified bit-for-bit against torch 2.11 CPU (tests/test_oracle). */
static inline float sq_norm3(const float *p) {
    float a = p[0] * p[0], b2 = p[1] * p[1], c = p[2] * p[2];
    return (a + b2) + c;
}
static inline float expansion_dist(const float *q, float qq, const float *c, float cc) {
    float dot = q[0] * c[0];
    dot = fmaf(q[1], c[1], dot);
    dot = fmaf(q[2], c[2], dot);
    float t = fmaf(-2.f, dot, qq);
    return t + cc;
}
void oracle_square_distance(int b, int s, int n, const float *src, const float *dst, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int i = 0; i < s; ++i) {
            const float *q = src + ((size_t)bi * s + i) * 3;
            float qq = sq_norm3(q);
            for (int k = 0; k < n; ++k) {
                const float *c = dst + ((size_t)bi * n + k) * 3;
                out[((size_t)bi * s + i) * n + k] = expansion_dist(q, qq, c, sq_norm3(c));
            }
        }
}

/* knn_point: pointconv_util.py:96-107.  torch.topk(sorted=False) leaves the
 * order unspecified and the choice among exact ties at the K-th boundary
 * unspecified; this oracle (and the CUDA kernel) fix both by ordering on
 * (distance, index).  Outputs are sorted ascending by that key. */
void oracle_knn(int b, int s, int n, int k, const float *query, const float *cand, int *idx, float *dist) {
    float *bd = (float *)malloc(sizeof(float) * k);
    int *bi_ = (int *)malloc(sizeof(int) * k);
    float *cc = (float *)malloc(sizeof(float) * n);
    for (int bi = 0; bi < b; ++bi) {
        const float *cb = cand + (size_t)bi * n * 3;
        for (int j = 0; j < n; ++j) cc[j] = sq_norm3(cb + j * 3);
        for (int i = 0; i < s; ++i) {
            const float *q = query + ((size_t)bi * s + i) * 3;
            float qq = sq_norm3(q);
            int cnt = 0;
            for (int j = 0; j < n; ++j) {
                float d = expansion_dist(q, qq, cb + j * 3, cc[j]);
                if (cnt == k && !(d < bd[k - 1])) continue;
                int pos = cnt < k ? cnt : k - 1;
                while (pos > 0 && d < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi_[pos] = bi_[pos - 1]; --pos; }
                bd[pos] = d; bi_[pos] = j;
                if (cnt < k) ++cnt;
            }
            for (int t = 0; t < k; ++t) {
                idx[((size_t)bi * s + i) * k + t] = t < cnt ? bi_[t] : 0;
                if (dist) dist[((size_t)bi * s + i) * k + t] = t < cnt ? bd[t] : INFINITY;
            }
        }
    }
    free(bd); free(bi_); free(cc);
}
