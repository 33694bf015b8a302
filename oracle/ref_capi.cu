// TEST INFRASTRUCTURE — not part of the product path.
//
// Thin C-ABI shim around the REFERENCE's own CUDA launchers, compiled byte-for-byte from
// /root/reference by oracle/build_ref.py into oracle/_ref/libref_ops.so (git-ignored build output;
// no reference source is copied into this repo).  It exists so that the GPU parity tests and the
// `ref_cuda` leg of bench.py can run the reference kernels (recompiled for sm_100a) on the same
// inputs as our kernels.  The reference's own .cpp bindings cannot be used: the six mm3d ones include
// the removed THC/THC.h (e.g. utils/mm3d_pn2/ops/ball_query/src/ball_query.cpp:4).
//
// Declarations below restate the launcher prototypes found at:
//   utils/mm3d_pn2/ops/furthest_point_sample/src/furthest_point_sample_cuda.cu:143,333
//   utils/mm3d_pn2/ops/ball_query/src/ball_query_cuda.cu:56
//   utils/mm3d_pn2/ops/gather_points/src/gather_points_cuda.cu:28,72
//   utils/mm3d_pn2/ops/group_points/src/group_points_cuda.cu:33,81
//   utils/mm3d_pn2/ops/interpolate/src/three_nn_cuda.cu:67
//   utils/mm3d_pn2/ops/interpolate/src/three_interpolate_cuda.cu:37,86
//   utils/mm3d_pn2/ops/knn/src/knn_cuda.cu:97
//   utils/metrics/CD/chamfer3D/chamfer3D.cu:136,176   (at::Tensor signatures)
//   utils/metrics/EMD/emd_cuda.cu:228,302             (at::Tensor signatures)
#include <ATen/ATen.h>
#include <cuda_runtime.h>

void furthest_point_sampling_kernel_launcher(int b, int n, int m, const float *dataset, float *temp,
                                             int *idxs, cudaStream_t stream);
void furthest_point_sampling_with_dist_kernel_launcher(int b, int n, int m, const float *dataset,
                                                       float *temp, int *idxs, cudaStream_t stream);
void ball_query_kernel_launcher(int b, int n, int m, float min_radius, float max_radius, int nsample,
                                const float *new_xyz, const float *xyz, int *idx, cudaStream_t stream);
void gather_points_kernel_launcher(int b, int c, int n, int npoints, const float *points,
                                   const int *idx, float *out, cudaStream_t stream);
void gather_points_grad_kernel_launcher(int b, int c, int n, int npoints, const float *grad_out,
                                        const int *idx, float *grad_points, cudaStream_t stream);
void group_points_kernel_launcher(int b, int c, int n, int npoints, int nsample, const float *points,
                                  const int *idx, float *out, cudaStream_t stream);
void group_points_grad_kernel_launcher(int b, int c, int n, int npoints, int nsample,
                                       const float *grad_out, const int *idx, float *grad_points,
                                       cudaStream_t stream);
void three_nn_kernel_launcher(int b, int n, int m, const float *unknown, const float *known,
                              float *dist2, int *idx, cudaStream_t stream);
void three_interpolate_kernel_launcher(int b, int c, int m, int n, const float *points, const int *idx,
                                       const float *weight, float *out, cudaStream_t stream);
void three_interpolate_grad_kernel_launcher(int b, int c, int n, int m, const float *grad_out,
                                            const int *idx, const float *weight, float *grad_points,
                                            cudaStream_t stream);
void knn_kernel_launcher(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz,
                         int *idx, float *dist2, cudaStream_t stream);

int chamfer_cuda_forward(at::Tensor xyz1, at::Tensor xyz2, at::Tensor dist1, at::Tensor dist2,
                         at::Tensor idx1, at::Tensor idx2);
int chamfer_cuda_backward(at::Tensor xyz1, at::Tensor xyz2, at::Tensor gradxyz1, at::Tensor gradxyz2,
                          at::Tensor graddist1, at::Tensor graddist2, at::Tensor idx1, at::Tensor idx2);
int emd_cuda_forward(at::Tensor xyz1, at::Tensor xyz2, at::Tensor dist, at::Tensor assignment,
                     at::Tensor price, at::Tensor assignment_inv, at::Tensor bid,
                     at::Tensor bid_increments, at::Tensor max_increments, at::Tensor unass_idx,
                     at::Tensor unass_cnt, at::Tensor unass_cnt_sum, at::Tensor cnt_tmp,
                     at::Tensor max_idx, float eps, int iters);
int emd_cuda_backward(at::Tensor xyz1, at::Tensor xyz2, at::Tensor gradxyz, at::Tensor graddist,
                      at::Tensor idx);

namespace {
at::Tensor f32(const void *p, std::initializer_list<int64_t> shape) {
  int dev = 0;
  cudaGetDevice(&dev);
  return at::from_blob(const_cast<void *>(p), shape,
                       at::TensorOptions().dtype(at::kFloat).device(at::kCUDA, dev));
}
at::Tensor i32(const void *p, std::initializer_list<int64_t> shape) {
  int dev = 0;
  cudaGetDevice(&dev);
  return at::from_blob(const_cast<void *>(p), shape,
                       at::TensorOptions().dtype(at::kInt).device(at::kCUDA, dev));
}
}  // namespace

extern "C" {

void ref_fps(int b, int n, int m, const float *xyz, float *temp, int *idx, void *stream) {
  furthest_point_sampling_kernel_launcher(b, n, m, xyz, temp, idx, (cudaStream_t)stream);
}
void ref_fps_with_dist(int b, int n, int m, const float *dist, float *temp, int *idx, void *stream) {
  furthest_point_sampling_with_dist_kernel_launcher(b, n, m, dist, temp, idx, (cudaStream_t)stream);
}
void ref_ball_query(int b, int n, int m, float min_radius, float max_radius, int nsample,
                    const float *new_xyz, const float *xyz, int *idx, void *stream) {
  ball_query_kernel_launcher(b, n, m, min_radius, max_radius, nsample, new_xyz, xyz, idx,
                             (cudaStream_t)stream);
}
void ref_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx,
                       float *out, void *stream) {
  gather_points_kernel_launcher(b, c, n, npoints, points, idx, out, (cudaStream_t)stream);
}
void ref_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                            float *grad_points, void *stream) {
  gather_points_grad_kernel_launcher(b, c, n, npoints, grad_out, idx, grad_points,
                                     (cudaStream_t)stream);
}
void ref_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                      const int *idx, float *out, void *stream) {
  group_points_kernel_launcher(b, c, n, npoints, nsample, points, idx, out, (cudaStream_t)stream);
}
void ref_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                           const int *idx, float *grad_points, void *stream) {
  group_points_grad_kernel_launcher(b, c, n, npoints, nsample, grad_out, idx, grad_points,
                                    (cudaStream_t)stream);
}
void ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                  int *idx, void *stream) {
  three_nn_kernel_launcher(b, n, m, unknown, known, dist2, idx, (cudaStream_t)stream);
}
void ref_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                           const float *weight, float *out, void *stream) {
  three_interpolate_kernel_launcher(b, c, m, n, points, idx, weight, out, (cudaStream_t)stream);
}
void ref_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                const float *weight, float *grad_points, void *stream) {
  three_interpolate_grad_kernel_launcher(b, c, n, m, grad_out, idx, weight, grad_points,
                                         (cudaStream_t)stream);
}
void ref_knn(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx,
             float *dist2, void *stream) {
  knn_kernel_launcher(b, n, m, nsample, xyz, new_xyz, idx, dist2, (cudaStream_t)stream);
}

// Chamfer / EMD launch on the legacy default stream in the reference (chamfer3D.cu:142, emd_cuda.cu:257).
int ref_chamfer_forward(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1,
                        float *dist2, int *idx1, int *idx2) {
  return chamfer_cuda_forward(f32(xyz1, {b, n, 3}), f32(xyz2, {b, m, 3}), f32(dist1, {b, n}),
                              f32(dist2, {b, m}), i32(idx1, {b, n}), i32(idx2, {b, m}));
}
int ref_chamfer_backward(int b, int n, int m, const float *xyz1, const float *xyz2, float *gradxyz1,
                         float *gradxyz2, const float *graddist1, const float *graddist2,
                         const int *idx1, const int *idx2) {
  return chamfer_cuda_backward(f32(xyz1, {b, n, 3}), f32(xyz2, {b, m, 3}), f32(gradxyz1, {b, n, 3}),
                               f32(gradxyz2, {b, m, 3}), f32(graddist1, {b, n}),
                               f32(graddist2, {b, m}), i32(idx1, {b, n}), i32(idx2, {b, m}));
}
// All 12 state arrays are caller-allocated and pre-initialised exactly as emd_module.py:54-65 does.
int ref_emd_forward(int b, int n, const float *xyz1, const float *xyz2, float *dist, int *assignment,
                    float *price, int *assignment_inv, int *bid, float *bid_increments,
                    float *max_increments, int *unass_idx, int *unass_cnt, int *unass_cnt_sum,
                    int *cnt_tmp, int *max_idx, float eps, int iters) {
  return emd_cuda_forward(f32(xyz1, {b, n, 3}), f32(xyz2, {b, n, 3}), f32(dist, {b, n}),
                          i32(assignment, {b, n}), f32(price, {b, n}), i32(assignment_inv, {b, n}),
                          i32(bid, {b, n}), f32(bid_increments, {b, n}), f32(max_increments, {b, n}),
                          i32(unass_idx, {(int64_t)b * n}), i32(unass_cnt, {512}),
                          i32(unass_cnt_sum, {512}), i32(cnt_tmp, {512}), i32(max_idx, {(int64_t)b * n}),
                          eps, iters);
}
int ref_emd_backward(int b, int n, const float *xyz1, const float *xyz2, float *gradxyz,
                     const float *graddist, const int *idx) {
  return emd_cuda_backward(f32(xyz1, {b, n, 3}), f32(xyz2, {b, n, 3}), f32(gradxyz, {b, n, 3}),
                           f32(graddist, {b, n}), i32(idx, {b, n}));
}

}  // extern "C"
