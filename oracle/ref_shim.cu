// oracle/ref_shim.cu -- TEST INFRASTRUCTURE.
// C-ABI doorway to the reference's own launchers (declared in /root/reference/lib/src/*_gpu.h,
// defined in the reference .cu files compiled next to this file by oracle/build_oracle.py --ref).
// Nothing here re-implements the reference; it only forwards raw pointers + stream.
#include <cuda_runtime.h>

void ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
                                     const float *new_xyz, const float *xyz, int *idx, cudaStream_t stream);
void group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                       const float *points, const int *idx, float *out, cudaStream_t stream);
void group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                            const float *grad_out, const int *idx, float *grad_points, cudaStream_t stream);
void gather_points_kernel_launcher_fast(int b, int c, int n, int npoints,
                                        const float *points, const int *idx, float *out, cudaStream_t stream);
void gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints,
                                             const float *grad_out, const int *idx, float *grad_points, cudaStream_t stream);
void furthest_point_sampling_kernel_launcher(int b, int n, int m,
                                             const float *dataset, float *temp, int *idxs, cudaStream_t stream);
void knn_kernel_launcher_fast(int b, int n, int m, int k, const float *unknown,
                              const float *known, float *dist2, int *idx, cudaStream_t stream);
void three_nn_kernel_launcher_fast(int b, int n, int m, const float *unknown,
                                   const float *known, float *dist2, int *idx, cudaStream_t stream);
void three_interpolate_kernel_launcher_fast(int b, int c, int m, int n,
                                            const float *points, const int *idx, const float *weight, float *out, cudaStream_t stream);
void three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m, const float *grad_out,
                                                 const int *idx, const float *weight, float *grad_points, cudaStream_t stream);

extern "C" {
void ref_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx, void *s) {
    ball_query_kernel_launcher_fast(b, n, m, radius, nsample, new_xyz, xyz, idx, (cudaStream_t)s); }
void ref_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out, void *s) {
    group_points_kernel_launcher_fast(b, c, n, npoints, nsample, points, idx, out, (cudaStream_t)s); }
void ref_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *g, const int *idx, float *gp, void *s) {
    group_points_grad_kernel_launcher_fast(b, c, n, npoints, nsample, g, idx, gp, (cudaStream_t)s); }
void ref_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out, void *s) {
    gather_points_kernel_launcher_fast(b, c, n, npoints, points, idx, out, (cudaStream_t)s); }
void ref_gather_points_grad(int b, int c, int n, int npoints, const float *g, const int *idx, float *gp, void *s) {
    gather_points_grad_kernel_launcher_fast(b, c, n, npoints, g, idx, gp, (cudaStream_t)s); }
void ref_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp, int *idxs, void *s) {
    furthest_point_sampling_kernel_launcher(b, n, m, dataset, temp, idxs, (cudaStream_t)s); }
void ref_knn(int b, int n, int m, int k, const float *unknown, const float *known, float *dist2, int *idx, void *s) {
    knn_kernel_launcher_fast(b, n, m, k, unknown, known, dist2, idx, (cudaStream_t)s); }
void ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, void *s) {
    three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx, (cudaStream_t)s); }
void ref_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *w, float *out, void *s) {
    three_interpolate_kernel_launcher_fast(b, c, m, n, points, idx, w, out, (cudaStream_t)s); }
void ref_three_interpolate_grad(int b, int c, int n, int m, const float *g, const int *idx, const float *w, float *gp, void *s) {
    three_interpolate_grad_kernel_launcher_fast(b, c, n, m, g, idx, w, gp, (cudaStream_t)s); }
}
