/* mvp_ops.h — C ABI of libmvp_ops.so: the B200 (sm_100a) point-cloud operator library that replaces
 * the native layer of paul007pl/MVP_Benchmark's `utils/metrics` and `utils/mm3d_pn2` packages.
 *
 * Conventions:
 *   - plain pointers and sizes only; no torch / pybind types.  Loadable with ctypes / dlopen.
 *   - every pointer is a DEVICE pointer on the current CUDA device unless its name ends in `_host`.
 *   - the CALLER allocates everything, including scratch (`*_workspace_bytes` tells how much);
 *     the callee never allocates, never synchronises the host, and launches on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - outputs need NO pre-initialisation: the zero-fills / sentinel-fills the reference's Python
 *     performs before each call (listed per function) are done inside.
 *   - all tensors are contiguous row-major fp32 / int32, shapes as in the reference.
 *   - return value: 0 = ok; > 0 = a cudaError_t; < 0 = one of MVP_ERR_* (invalid argument).
 *     The reference printf()s / exit(-1)s instead (chamfer3D.cu:145-150, ball_query_cuda.cu:73-77);
 *     the Python layer here raises RuntimeError on any non-zero code.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference root).
 */
#ifndef MVP_OPS_H_
#define MVP_OPS_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVP_OK 0
#define MVP_ERR_INVALID_ARGUMENT (-1)  /* negative size, null pointer, k out of range ...          */
#define MVP_ERR_EMD_SIZE_MISMATCH (-2) /* emd_cuda.cu:236-239 "two point clouds should have the same size" */
#define MVP_ERR_EMD_BATCH (-3)         /* emd_cuda.cu:241-244 batch size > 512                      */
#define MVP_ERR_EMD_MULTIPLE_1024 (-4) /* emd_cuda.cu:246-249 n % 1024 != 0                         */
#define MVP_ERR_WORKSPACE (-5)         /* workspace pointer null or too small                      */
#define MVP_ERR_UNSUPPORTED_DEVICE (-6)/* not an sm_100 device                                     */

typedef void *mvp_stream_t; /* cudaStream_t */

/* Library / build identification. */
int mvp_abi_version(void);
const char *mvp_build_info(void);
/* Text for a return code of any function below (MVP_ERR_* or cudaError_t). */
const char *mvp_error_string(int code);
/* Number of KERNELS (memsets excluded) this library has launched in this process; monotonic. */
unsigned long long mvp_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Chamfer distance.
 * Replaces chamfer_3D.forward  -> chamfer_cuda_forward  (utils/metrics/CD/chamfer3D/chamfer_cuda.cpp:17-19,
 *          utils/metrics/CD/chamfer3D/chamfer3D.cu:136-154, kernel :12-134).
 *   xyz1 (b,n,3), xyz2 (b,m,3) -> dist1 (b,n), idx1 (b,n): min_j |xyz1_i - xyz2_j|^2 and its argmin
 *   (lowest j on ties); dist2 (b,m), idx2 (b,m): the same with roles swapped.  Bit-exact with the
 *   reference kernel for finite inputs.  Outputs need not be zeroed (dist_chamfer_3D.py:33-42 does).
 *   workspace: mvp_chamfer_forward_workspace_bytes(b,n,m) bytes, 16-byte aligned.
 * ------------------------------------------------------------------------------------------- */
size_t mvp_chamfer_forward_workspace_bytes(int b, int n, int m);
int mvp_chamfer_forward(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1,
                        float *dist2, int *idx1, int *idx2, void *workspace, size_t workspace_bytes,
                        mvp_stream_t stream);

/* The same operator with the algorithm named explicitly (tests and benchmarks use it; results are
 * bit-identical for all three):
 *   MVP_CHAMFER_AUTO  what mvp_chamfer_forward does: GRID when the shape supports it, else BRUTE;
 *   MVP_CHAMFER_BRUTE every pair evaluated (tiled B*N*M argmin, both directions from one evaluation);
 *   MVP_CHAMFER_GRID  exact nearest neighbour through a uniform grid built per cloud, pruned with
 *                     conservative lower bounds; what the search does not finish (far / very dense
 *                     surroundings, degenerate grids) is completed by a warp-cooperative pass over the rows
 *                     of the grid (n, m >= 512; MVP_ERR_INVALID_ARGUMENT otherwise). */
#define MVP_CHAMFER_AUTO 0
#define MVP_CHAMFER_BRUTE 1
#define MVP_CHAMFER_GRID 2
int mvp_chamfer_forward_algo(int algo, int b, int n, int m, const float *xyz1, const float *xyz2,
                             float *dist1, float *dist2, int *idx1, int *idx2, void *workspace,
                             size_t workspace_bytes, mvp_stream_t stream);

/* Replaces chamfer_3D.backward -> chamfer_cuda_backward (chamfer_cuda.cpp:22-26, chamfer3D.cu:176-195,
 * kernel :155-174).  gradxyz1 (b,n,3) and gradxyz2 (b,m,3) need no zero-fill (dist_chamfer_3D.py:56-57 does
 * it in the reference): each point's own half is written with a plain store, the scattered halves are
 * accumulated with fp32 atomics. */
int mvp_chamfer_backward(int b, int n, int m, const float *xyz1, const float *xyz2,
                         const float *graddist1, const float *graddist2, const int *idx1,
                         const int *idx2, float *gradxyz1, float *gradxyz2, mvp_stream_t stream);

/* The same with the algorithm named explicitly:
 *   MVP_CHAMFER_BWD_AUTO    what mvp_chamfer_backward does: own halves written, scattered halves added with vector
 *                           reductions (red.global.add.v4/v2.f32) — summation order not deterministic, like the
 *                           reference's atomics;
 *   MVP_CHAMFER_BWD_SUMMED  no float atomics: the index is transposed in shared memory and every gradient row is
 *                           summed by one thread, lists of up to 8 contributions in ascending source order
 *                           (n, m <= 16384; MVP_ERR_INVALID_ARGUMENT otherwise). */
#define MVP_CHAMFER_BWD_AUTO 0
#define MVP_CHAMFER_BWD_SUMMED 1
int mvp_chamfer_backward_algo(int algo, int b, int n, int m, const float *xyz1, const float *xyz2,
                              const float *graddist1, const float *graddist2, const int *idx1,
                              const int *idx2, float *gradxyz1, float *gradxyz2, mvp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Earth mover's distance, auction approximation.
 * Replaces emd.forward -> emd_cuda_forward (utils/metrics/EMD/emd.cpp:12-16, emd_cuda.cu:228-282; kernels
 * :23-226).  xyz1, xyz2 (b,n,3), n % 1024 == 0, b <= 512 -> dist (b,n), assignment (b,n).
 * The twelve state arrays the reference's Python allocates (emd_module.py:54-65) live in `workspace`
 * and are initialised inside.  Near-tie winners (emd_cuda.cu:188, a last-writer race in the reference)
 * are resolved deterministically: the highest source index wins.
 * ------------------------------------------------------------------------------------------- */
size_t mvp_emd_forward_workspace_bytes(int b, int n);
int mvp_emd_forward(int b, int n, int m, const float *xyz1, const float *xyz2, float eps, int iters,
                    float *dist, int *assignment, void *workspace, size_t workspace_bytes,
                    mvp_stream_t stream);

/* The same operator with the Bid search named explicitly (tests and benchmarks; results are bit-identical):
 *   MVP_EMD_AUTO   what mvp_emd_forward does: GRID when n <= 8192, else BRUTE;
 *   MVP_EMD_BRUTE  every unassigned source scans all n targets each round (as the reference's Bid kernel does);
 *   MVP_EMD_GRID   sources search a uniform grid over the targets outwards and stop when no unvisited target can
 *                  reach their second-best value (n <= 8192; MVP_ERR_INVALID_ARGUMENT otherwise). */
#define MVP_EMD_AUTO 0
#define MVP_EMD_BRUTE 1
#define MVP_EMD_GRID 2
int mvp_emd_forward_algo(int algo, int b, int n, int m, const float *xyz1, const float *xyz2, float eps,
                         int iters, float *dist, int *assignment, void *workspace, size_t workspace_bytes,
                         mvp_stream_t stream);

/* Replaces emd.backward -> emd_cuda_backward (emd.cpp:18-21, emd_cuda.cu:302-316, kernel :284-300).
 * gradxyz1 (b,n,3) is fully written (no pre-zeroing needed).  The reference returns zeros for xyz2
 * (emd_module.py:78-81); that tensor is the Python layer's business. */
int mvp_emd_backward(int b, int n, const float *xyz1, const float *xyz2, const float *graddist,
                     const int *assignment, float *gradxyz1, mvp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * PointNet++ ops (utils/mm3d_pn2/ops).
 * ------------------------------------------------------------------------------------------- */

/* Replaces furthest_point_sampling_wrapper -> furthest_point_sampling_kernel_launcher
 * (furthest_point_sample/src/furthest_point_sample.cpp:32-43, furthest_point_sample_cuda.cu:143-209,
 * kernel :26-141).  xyz (b,n,3) -> idx (b,m).  `temp` (b,n) is the reference's scratch of running
 * minimum distances (pre-filled with 1e10 by furthest_point_sample.py:30); here it may be NULL — when
 * given, the final running distances are written to it, no pre-fill needed.  Bit-exact incl. ties. */
int mvp_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                mvp_stream_t stream);

/* Replaces furthest_point_sampling_with_dist_wrapper (furthest_point_sample.cpp:45-57,
 * furthest_point_sample_cuda.cu:333-399, kernel :214-331).  dist (b,n,n) -> idx (b,m). */
int mvp_furthest_point_sampling_with_dist(int b, int n, int m, const float *dist, float *temp, int *idx,
                                          mvp_stream_t stream);

/* Replaces ball_query_wrapper -> ball_query_kernel_launcher (ball_query/src/ball_query.cpp:30-43,
 * ball_query_cuda.cu:56-78, kernel :11-54).  new_xyz (b,m,3) centres, xyz (b,n,3) -> idx (b,m,nsample);
 * rows without any hit are zero (ball_query.py:35 zeroes idx in the reference; done inside here). */
int mvp_ball_query(int b, int n, int m, float min_radius, float max_radius, int nsample,
                   const float *new_xyz, const float *xyz, int *idx, mvp_stream_t stream);

/* Replaces gather_points_wrapper / gather_points_grad_wrapper (gather_points/src/gather_points.cpp:28-52,
 * gather_points_cuda.cu:28-44,72-90).  points (b,c,n), idx (b,npoints) -> out (b,c,npoints);
 * grad_out (b,c,npoints) -> grad_points (b,c,n), zero-filled inside (gather_points.py:44). */
int mvp_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out,
                      mvp_stream_t stream);
int mvp_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                           float *grad_points, mvp_stream_t stream);

/* Replaces group_points_ext.forward / backward (group_points/src/group_points.cpp:31-57,
 * group_points_cuda.cu:81-98,33-51).  points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample);
 * grad_points zero-filled inside (group_points.py:213). */
int mvp_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                     float *out, mvp_stream_t stream);
int mvp_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int *idx, float *grad_points, mvp_stream_t stream);

/* Replaces three_nn_wrapper (interpolate/src/interpolate.cpp:46-56, three_nn_cuda.cu:67-86, kernel :11-65).
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) SQUARED distances ascending, idx (b,n,3). */
int mvp_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                 mvp_stream_t stream);

/* Replaces three_interpolate_wrapper / _grad_wrapper (interpolate.cpp:58-85,
 * three_interpolate_cuda.cu:37-59,86-106).  points (b,c,m), idx (b,n,3), weight (b,n,3) -> out (b,c,n);
 * grad_out (b,c,n) -> grad_points (b,c,m), zero-filled inside (three_interpolate.py:53). */
int mvp_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                          const float *weight, float *out, mvp_stream_t stream);
int mvp_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                               const float *weight, float *grad_points, mvp_stream_t stream);

/* Replaces knn_wrapper (knn/src/knn.cpp:28-41, knn_cuda.cu:97-115, kernel :58-94).
 * xyz (b,n,3), new_xyz (b,m,3) centres, 0 < nsample <= 100 -> idx (b,m,nsample), dist2 (b,m,nsample),
 * ascending distance.  (The transpose to (b,nsample,m) stays in Python, knn.py:63.) */
int mvp_knn(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx,
            float *dist2, mvp_stream_t stream);

/* The F-score of a Chamfer result (utils/metrics/CD/fscore.py:12-15; completion/model_utils.py:74-76 with calc_f1):
 * precision_k[b] = mean_i (dist_k[b,i] < threshold), fscore = 2 p1 p2 / (p1 + p2) with NaN -> 0, torch's arithmetic
 * (mean = count * fl(1/n)).  dist1 (b,n), dist2 (b,m); outputs (b). */
int mvp_fscore(int b, int n, int m, const float *dist1, const float *dist2, float threshold, float *fscore,
               float *precision1, float *precision2, mvp_stream_t stream);

/* SURVEY.md §8(f) row 3 — the loss epilogue the models compute from the Chamfer outputs with torch glue
 * (completion/model_utils.py:67-72, calc_cd):  cd_p[b] = (mean_i sqrt(dist1[b,i]) + mean_j sqrt(dist2[b,j])) / 2,
 * cd_t[b] = mean_i dist1[b,i] + mean_j dist2[b,j];  dist1 (b,n), dist2 (b,m) -> cd_p (b), cd_t (b).  fp32, fixed
 * summation order (deterministic); within 1e-5 relative of torch's reductions.  The gradient entry point returns
 * grad_dist1 (b,n), grad_dist2 (b,m) from the upstream gradients of cd_p and cd_t (b each):
 * grad_cd_p / (4 n sqrt(d)) + grad_cd_t / n — infinite at d == 0, as torch's sqrt backward.  Opt-in
 * (mvp_benchmark_b200.model_patches rebinds calc_cd). */
int mvp_chamfer_loss(int b, int n, int m, const float *dist1, const float *dist2, float *cd_p, float *cd_t,
                     mvp_stream_t stream);
int mvp_chamfer_loss_grad(int b, int n, int m, const float *dist1, const float *dist2, const float *grad_cd_p,
                          const float *grad_cd_t, float *grad_dist1, float *grad_dist2, mvp_stream_t stream);

/* SURVEY.md §8(f) row 2 — the glue between three_nn and three_interpolate (completion/model_utils.py:286-293,
 * three_nn_upsampling): dist2 (b,n,3) SQUARED distances as mvp_three_nn returns them -> weight (b,n,3),
 * w_k = (1/d_k) / (1/d_0 + 1/d_1 + 1/d_2) with d_k = max(sqrt(dist2_k), 1e-10): the sqrt of three_nn.py:38 and five
 * torch kernels in one pass, IEEE operations (the sum in the order torch 2.11 adds three elements; within 2 ulp of
 * any other order).  Opt-in (model_patches rebinds three_nn_upsampling). */
int mvp_three_nn_weights(int b, int n, const float *dist2, float *weight, mvp_stream_t stream);

/* three_nn (workspace variant above) AND those weights from the same launches: the kernel that finds a target's three
 * neighbours writes its weights too.  Outputs dist2, idx, weight, each (b,n,3). */
int mvp_three_nn_weights_ws(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                            float *weight, void *workspace, size_t workspace_bytes, mvp_stream_t stream);

/* gather_points over the pk neighbours of every sampled point followed by torch.max over the neighbours
 * (completion/model_utils.py:97-102: edge_preserve_sampling) as ONE launch: points (b,c,n), idx (b,npoints,k) ->
 * out (b,c,npoints) and arg (b,c,npoints) int32, the neighbour index attaining each maximum (first among equals).
 * n <= 16384.  mvp_gather_max_grad: grad_points (b,c,n) = grad_out scattered by arg, fully written. */
int mvp_gather_max(int b, int c, int n, int npoints, int k, const float *points, const int *idx, float *out, int *arg,
                   mvp_stream_t stream);
int mvp_gather_max_grad(int b, int c, int n, int npoints, const float *grad_out, const int *arg, float *grad_points,
                        mvp_stream_t stream);

/* SA_module's attention-weighted aggregation of the neighbours (completion/models/vrcnet.py:49-52: get_edge_features,
 * repeat of the weights over share_planes, multiply, sum over k) as ONE launch without its (B, C, k, N) intermediates:
 * out[b,ch,p] = sum_j w[b, ch mod cw, j, p] * y[b, ch, idx[b,p,j]].  y (b,c,n), idx (b,n,k) int32, w (b,cw,k,n),
 * out (b,c,n); c = S * cw with S <= 8 channels sharing a weight row; n <= 6144.  _grad: grad_y (b,c,n) and
 * grad_w (b,cw,k,n), fully written. */
int mvp_neighbor_weighted_sum(int b, int c, int cw, int n, int k, const float *y, const int *idx, const float *w,
                              float *out, mvp_stream_t stream);
int mvp_neighbor_weighted_sum_grad(int b, int c, int cw, int n, int k, const float *y, const int *idx, const float *w,
                                   const float *grad_out, float *grad_y, float *grad_w, mvp_stream_t stream);

/* A shared 1x1 convolution over a cloud's feature map on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulate in
 * tensor memory): y[b,co,p] = sum_ci w[co,ci] * x[b,ci,p] (+ bias[co]) (then max(., 0) if relu != 0) — what the
 * nn.Conv1d / nn.Conv2d(kernel_size=1) layers of completion/models/{pcn,ecg,vrcnet}.py and of
 * completion/model_utils.py compute (cuDNN runs them in TF32 too: torch.backends.cudnn.allow_tf32 defaults to True).
 * x (b,cin,n), w (cout,cin), bias (cout) or NULL, y (b,cout,n), all fp32; operands are rounded to TF32 (round to
 * nearest) as they are staged.  b <= 65535.
 * mvp_pointwise_conv_masked: the same contraction of x * (mask > 0), mask (b,cin,n), no bias — the input gradient of a
 * layer followed by ReLU: x = the gradient of the ReLU's output, mask = that output, w = the layer's weight transposed. */
int mvp_pointwise_conv(int b, int cin, int cout, int n, const float *x, const float *w, const float *bias, int relu,
                       float *y, mvp_stream_t stream);
int mvp_pointwise_conv_masked(int b, int cin, int cout, int n, const float *x, const float *mask, const float *w, float *y,
                              mvp_stream_t stream);

/* The weight and bias gradients of the same layer, again on the tensor cores (UMMA M = output channels, N = input
 * channels, K = points: both operands K-major as they lie): grad_w[o,c] = sum_b sum_p g[b,o,p] * x[b,c,p], (cout,cin);
 * grad_bias[o] = sum_b sum_p g[b,o,p] (a row of ones appended to x in shared memory) or NULL.  g (b,cout,n), x (b,cin,n).
 * cin + (grad_bias != NULL) <= 256.  Deterministic: persistent CTAs write partial products, added in a fixed order.
 * workspace: mvp_pointwise_wgrad_workspace_bytes(cin, cout, with_bias) bytes of device memory. */
size_t mvp_pointwise_wgrad_workspace_bytes(int cin, int cout, int with_bias);
int mvp_pointwise_wgrad(int b, int cin, int cout, int n, const float *g, const float *x, float *grad_w, float *grad_bias,
                        void *workspace, size_t workspace_bytes, mvp_stream_t stream);

/* The bias of a wide 1x1 layer that stays on a library GEMM: y[b,c,:] += bias[c] in place (then max(., 0) if relu),
 * y (b,c,n); and its gradient: out[c] = sum over clouds and points of g[b,c,p], g (b,c,n), in a fixed order
 * (deterministic).  workspace: mvp_channel_sum_workspace_bytes(b, c) bytes of device memory. */
int mvp_bias_add(int b, int c, int n, float *y, const float *bias, int relu, mvp_stream_t stream);
size_t mvp_channel_sum_workspace_bytes(int b, int c);
int mvp_channel_sum(int b, int c, int n, const float *g, float *out, void *workspace, size_t workspace_bytes,
                    mvp_stream_t stream);

/* The maximum over a point's k neighbours — the last axis of a contiguous (rows, k) fp32 view, k <= 255: what
 * `y, _ = torch.max(y, 3)` computes on the (B, C, N, k) neighbour tensors of completion/models/ecg.py:64 and
 * completion/model_utils.py:53,104.  out (rows), arg (rows) uint8 = the first position of the maximum (a NaN wins, the
 * first one, as in torch).  mvp_max_last_grad: grad_x (rows, k) = grad_out at arg, zero elsewhere, fully written. */
int mvp_max_last(long long rows, int k, const float *x, float *out, unsigned char *arg, mvp_stream_t stream);
int mvp_max_last_grad(long long rows, int k, const float *grad_out, const unsigned char *arg, float *grad_x,
                      mvp_stream_t stream);

/* The k <= 32 largest entries of every row of a (rows, cols) fp32 score matrix, descending, equal scores in ascending
 * column order — what completion/model_utils.py:242-247 asks torch.topk for on its (B, N, N) matrix of negative
 * feature-space distances.  Any of values (rows,k) / idx64 (rows,k) int64 / idx32 (rows,k) int32 may be NULL. */
int mvp_topk_rows(long long rows, int cols, int k, const float *scores, float *values, long long *idx64, int *idx32,
                  mvp_stream_t stream);

/* The same selection on the feature-space kNN score of completion/model_utils.py:242-247 WITHOUT its (B, N, N) score
 * matrix: gram (b,n,n) = x^T x (the original's torch.matmul), sqnorm (b,n) = sum_c x^2 (its torch.sum); the ranked
 * value of row i, column j is  (-sqnorm[j] - (-2 * gram[i,j])) - sqnorm[i], computed with the IEEE operations torch's
 * three elementwise kernels perform, in their order: identical values and indices. */
int mvp_topk_rows_sqdist(int b, int n, int k, const float *gram, const float *sqnorm, float *values, long long *idx64,
                         int *idx32, mvp_stream_t stream);

/* furthest_point_sample followed by gather_points on the transposed cloud (completion/model_utils.py:91-93,
 * completion/models/vrcnet.py:451) as ONE launch: idx (b,m) as mvp_furthest_point_sampling, and the sampled points
 * themselves, (b,m,3) or — channels_first != 0 — (b,3,m), the layout gather_points returns. */
int mvp_furthest_point_sampling_gather(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                       float *sampled_xyz, int channels_first, mvp_stream_t stream);

/* ball_query followed by grouping_operation on the coordinates and permute(0,2,3,1) (completion/model_utils.py:211-214)
 * as ONE launch: idx (b,m,nsample) as mvp_ball_query, and the neighbours' coordinates (b,m,nsample,3). */
int mvp_ball_query_group(int b, int n, int m, float min_radius, float max_radius, int nsample, const float *new_xyz,
                         const float *xyz, int *idx, float *grouped_xyz, mvp_stream_t stream);

/* SURVEY.md §8(f) row 1 — the k-nearest-neighbour search the completion MODELS run in torch
 * (completion/model_utils.py:242-259 `knn` / `knn_point` / `knn_point_all`: a (B,N,M) matrix of
 * -|x|^2 + 2 x.y - |y|^2 by matmul, then torch.topk), as one fused exact search for 3-D points.
 * Opt-in (the call sites are caller code): mvp_benchmark_b200.model_patches replaces the three functions.
 *   queries (b,n,3), cloud (b,m,3), 1 <= k <= min(m, 64)  ->  dist2 (b,n,k) squared distances, idx (b,n,k),
 *   ascending in (distance, index): distances are fl(dx*dx + dy*dy + dz*dz) in the contraction every other
 *   operator here uses, equal distances resolve to the lower index.  (The matmul expansion the models use
 *   rounds differently — its distances differ by ~1e-7 and may order near-ties the other way; its order among
 *   exact ties is unspecified.)  Exact: grid-pruned search for k <= 32 and n, m >= 256, exhaustive otherwise;
 *   both give the same bits.  workspace: mvp_knn_points_workspace_bytes(b,n,m,k) bytes, 16-byte aligned. */
size_t mvp_knn_points_workspace_bytes(int b, int n, int m, int k);
int mvp_knn_points(int b, int n, int m, int k, const float *queries, const float *cloud, float *dist2,
                   int *idx, void *workspace, size_t workspace_bytes, mvp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The backward scatters again, with a caller-provided workspace (same results, faster): the index is
 * transposed once per cloud into the workspace (start[rows + 1] | perm[entries] | weights permuted alike) and every gradient
 * element is then SUMMED by one thread and written once — no floating-point atomics, no memset.
 *   rows    = destination columns per channel (n of gather / group, m of three_interpolate)
 *   entries = index entries per cloud (npoints, npoints*nsample, 3*n of three_interpolate)
 * mvp_scatter_workspace_bytes(b, rows, entries) bytes, 16-byte aligned.  Shapes for which this path
 * is not the faster one (rows of grad_out above 32 KB, fewer entries than destinations, rows > 49152)
 * and a missing / short workspace fall back to the functions above.
 * ------------------------------------------------------------------------------------------- */
size_t mvp_scatter_workspace_bytes(int b, int rows, int entries);
int mvp_gather_points_grad_ws(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                              float *grad_points, void *workspace, size_t workspace_bytes,
                              mvp_stream_t stream);
int mvp_group_points_grad_ws(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                             const int *idx, float *grad_points, void *workspace,
                             size_t workspace_bytes, mvp_stream_t stream);
int mvp_three_interpolate_grad_ws(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                  const float *weight, float *grad_points, void *workspace,
                                  size_t workspace_bytes, mvp_stream_t stream);

/* three_nn with a workspace: the three nearest sources are found through a uniform grid over `known`
 * (the search of the grid-pruned Chamfer path with a top-3 state; targets it does not finish go through
 * the exhaustive kernel) — same outputs as mvp_three_nn, bit for bit.  n, m >= 256; other shapes and a
 * missing / short workspace fall back to mvp_three_nn.  mvp_three_nn_workspace_bytes(b, n, m) bytes. */
size_t mvp_three_nn_workspace_bytes(int b, int n, int m);
int mvp_three_nn_ws(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                    void *workspace, size_t workspace_bytes, mvp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MVP_OPS_H_ */
