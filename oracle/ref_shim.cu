// TEST INFRASTRUCTURE ONLY — not product code.
//
// extern "C" shim over the reference's own CUDA launchers, so that tests on the
// GPU box can call the UNMODIFIED reference kernels through ctypes with raw
// device pointers.  The launchers themselves are compiled from where they lie
// under /root/reference/pointnet2/src (see oracle/Makefile); nothing from the
// reference is copied into this repository.  The launchers are declared in
//   pointnet2/src/sampling_gpu.h:12-27, group_points_gpu.h, interpolate_gpu.h,
//   ball_query_gpu.h
// with C++ linkage; this file only re-exports them with C linkage.
#include <cuda_runtime_api.h>

void gather_points_kernel_launcher_fast(int b, int c, int n, int npoints,
    const float *points, const int *idx, float *out, cudaStream_t stream);
void gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints,
    const float *grad_out, const int *idx, float *grad_points, cudaStream_t stream);
void furthest_point_sampling_kernel_launcher(int b, int n, int m,
    const float *dataset, float *temp, int *idxs, cudaStream_t stream);
void group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
    const float *points, const int *idx, float *out, cudaStream_t stream);
void group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
    const float *grad_out, const int *idx, float *grad_points, cudaStream_t stream);
void three_nn_kernel_launcher_fast(int b, int n, int m, const float *unknown,
    const float *known, float *dist2, int *idx, cudaStream_t stream);
void three_interpolate_kernel_launcher_fast(int b, int c, int m, int n,
    const float *points, const int *idx, const float *weight, float *out, cudaStream_t stream);
void three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m,
    const float *grad_out, const int *idx, const float *weight, float *grad_points, cudaStream_t stream);
void ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
    const float *new_xyz, const float *xyz, int *idx, cudaStream_t stream);

extern "C" {

void ref_fps(int b, int n, int m, const float *xyz, float *temp, int *idx, cudaStream_t s) {
    furthest_point_sampling_kernel_launcher(b, n, m, xyz, temp, idx, s);
}
void ref_gather(int b, int c, int n, int m, const float *f, const int *idx, float *out, cudaStream_t s) {
    gather_points_kernel_launcher_fast(b, c, n, m, f, idx, out, s);
}
void ref_gather_grad(int b, int c, int n, int m, const float *g, const int *idx, float *gf, cudaStream_t s) {
    gather_points_grad_kernel_launcher_fast(b, c, n, m, g, idx, gf, s);
}
void ref_group(int b, int c, int n, int np, int ns, const float *f, const int *idx, float *out, cudaStream_t s) {
    group_points_kernel_launcher_fast(b, c, n, np, ns, f, idx, out, s);
}
void ref_group_grad(int b, int c, int n, int np, int ns, const float *g, const int *idx, float *gf, cudaStream_t s) {
    group_points_grad_kernel_launcher_fast(b, c, n, np, ns, g, idx, gf, s);
}
void ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, cudaStream_t s) {
    three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx, s);
}
void ref_three_interpolate(int b, int c, int m, int n, const float *f, const int *idx, const float *w, float *out, cudaStream_t s) {
    three_interpolate_kernel_launcher_fast(b, c, m, n, f, idx, w, out, s);
}
void ref_three_interpolate_grad(int b, int c, int n, int m, const float *g, const int *idx, const float *w, float *gf, cudaStream_t s) {
    three_interpolate_grad_kernel_launcher_fast(b, c, n, m, g, idx, w, gf, s);
}
void ref_ball_query(int b, int n, int m, float radius, int ns, const float *new_xyz, const float *xyz, int *idx, cudaStream_t s) {
    ball_query_kernel_launcher_fast(b, n, m, radius, ns, new_xyz, xyz, idx, s);
}

}  // extern "C"
