// Loss epilogue of the Chamfer distance — SURVEY.md §8(f) row 3.
//
// The completion models turn the four outputs of the Chamfer operator into two numbers per cloud with torch glue
// (completion/model_utils.py:67-77, calc_cd):
//     cd_p = (sqrt(dist1).mean(1) + sqrt(dist2).mean(1)) / 2        cd_t = dist1.mean(1) + dist2.mean(1)
// i.e. two sqrt kernels, four reductions and three elementwise kernels per call, four calls per VRCNet step.  Here:
// one CTA per cloud sums sqrt(d) and d over both directions in one pass (fp32, fixed order: strided per-thread
// partial sums, shuffle tree, 8 warp partials summed by thread 0 — deterministic, unlike a float atomic), and one
// elementwise kernel produces both upstream gradients  d cd_p / d dist = 1 / (4 n sqrt(d)),  d cd_t / d dist = 1 / n
// (infinite at d == 0, exactly like torch's sqrt backward).  Opt-in: mvp_benchmark_b200.model_patches rebinds calc_cd.
#include "common.cuh"

namespace mvp {

constexpr int kLossThreads = 256;

__global__ void __launch_bounds__(kLossThreads)
chamfer_loss_kernel(int n, int m, const float *__restrict__ dist1, const float *__restrict__ dist2,
                    float *__restrict__ cd_p, float *__restrict__ cd_t) {
  __shared__ float s_part[4][kLossThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *d1 = dist1 + (size_t)b * n, *d2 = dist2 + (size_t)b * m;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // sum sqrt(d1), sum d1, sum sqrt(d2), sum d2
  for (int i = tid; i < n; i += kLossThreads) {
    const float d = __ldg(d1 + i);
    acc[0] += __fsqrt_rn(d);
    acc[1] += d;
  }
  for (int i = tid; i < m; i += kLossThreads) {
    const float d = __ldg(d2 + i);
    acc[2] += __fsqrt_rn(d);
    acc[3] += d;
  }
#pragma unroll
  for (int off = 16; off; off >>= 1)
#pragma unroll
    for (int a = 0; a < 4; a++) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], off);
  if (lane == 0)
#pragma unroll
    for (int a = 0; a < 4; a++) s_part[a][warp] = acc[a];
  __syncthreads();
  if (tid == 0) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int w = 0; w < kLossThreads / 32; w++)
#pragma unroll
      for (int a = 0; a < 4; a++) t[a] += s_part[a][w];
    cd_p[b] = (t[0] / (float)n + t[2] / (float)m) / 2;
    cd_t[b] = t[1] / (float)n + t[3] / (float)m;
  }
}

__global__ void __launch_bounds__(256)
chamfer_loss_grad_kernel(int b, int n, int m, const float *__restrict__ dist1, const float *__restrict__ dist2,
                         const float *__restrict__ g_p, const float *__restrict__ g_t, float *__restrict__ gd1,
                         float *__restrict__ gd2) {
  const long long total1 = (long long)b * n, total = total1 + (long long)b * m;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const bool first = i < total1;
    const long long p = first ? i : i - total1;
    const int cnt = first ? n : m;
    const int cloud = (int)(p / cnt);
    const float d = __ldg((first ? dist1 : dist2) + p);
    // torch: mean backward (g / cnt), then sqrt backward (g / (2 sqrt(d))), after the "/ 2" of cd_p
    const float gp = (__ldg(g_p + cloud) / 2) / (float)cnt;
    const float gt = __ldg(g_t + cloud) / (float)cnt;
    // an absent / zero upstream gradient of cd_p contributes nothing: autograd never runs sqrt's backward for an
    // unused output, so d == 0 must not turn 0 / 0 into NaN here (cd_t-only training with coincident points)
    (first ? gd1 : gd2)[p] = (gp != 0.f ? gp / (2 * __fsqrt_rn(d)) : 0.f) + gt;
  }
}

// F-score of a Chamfer result (utils/metrics/CD/fscore.py:12-15): precision_k = mean over the cloud of (dist_k <
// threshold), fscore = 2 p1 p2 / (p1 + p2), NaN (p1 = p2 = 0) -> 0 — four elementwise kernels, two reductions and an
// indexed assignment in torch; here one CTA per cloud counts both directions.  The arithmetic is torch's: a mean is
// count * fl(1 / n) (its reduce kernel multiplies the sum by a float factor), then ((2 p1) p2) / (p1 + p2).
__global__ void __launch_bounds__(kLossThreads)
fscore_kernel(int n, int m, float threshold, const float *__restrict__ dist1, const float *__restrict__ dist2,
              float *__restrict__ f, float *__restrict__ p1, float *__restrict__ p2) {
  __shared__ int s_part[2][kLossThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *d1 = dist1 + (size_t)b * n, *d2 = dist2 + (size_t)b * m;
  int c1 = 0, c2 = 0;
  for (int i = tid; i < n; i += kLossThreads) c1 += __ldg(d1 + i) < threshold ? 1 : 0;
  for (int i = tid; i < m; i += kLossThreads) c2 += __ldg(d2 + i) < threshold ? 1 : 0;
  c1 = __reduce_add_sync(0xffffffffu, c1);
  c2 = __reduce_add_sync(0xffffffffu, c2);
  if (lane == 0) s_part[0][warp] = c1, s_part[1][warp] = c2;
  __syncthreads();
  if (tid == 0) {
    int t1 = 0, t2 = 0;
    for (int w = 0; w < kLossThreads / 32; w++) t1 += s_part[0][w], t2 += s_part[1][w];
    const float q1 = __fmul_rn((float)t1, __fdiv_rn(1.f, (float)n)), q2 = __fmul_rn((float)t2, __fdiv_rn(1.f, (float)m));
    const float v = __fdiv_rn(__fmul_rn(__fmul_rn(2.f, q1), q2), __fadd_rn(q1, q2));
    f[b] = v != v ? 0.f : v;
    p1[b] = q1;
    p2[b] = q2;
  }
}

// Inverse-distance weights of the three nearest sources (completion/model_utils.py:286-293, three_nn_upsampling):
//     dist = max(sqrt(d2), 1e-10);  w_k = (1 / dist_k) / ((1 / dist_0 + 1 / dist_2) + 1 / dist_1)
// from the SQUARED distances mvp_three_nn returns — the sqrt of three_nn.py:38 and the five torch kernels of the
// glue in one elementwise pass, every operation the IEEE one torch performs, in torch's order.
__global__ void __launch_bounds__(256)
three_nn_weights_kernel(long long total, const float *__restrict__ dist2, float *__restrict__ weight) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    three_nn_weights_of(__ldg(dist2 + i * 3), __ldg(dist2 + i * 3 + 1), __ldg(dist2 + i * 3 + 2), weight + i * 3);
  }
}

}  // namespace mvp

using namespace mvp;

MVP_API int mvp_three_nn_weights(int b, int n, const float *dist2, float *weight, mvp_stream_t stream) {
  if (b < 0 || n < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0) return MVP_OK;
  if (!dist2 || !weight) return MVP_ERR_INVALID_ARGUMENT;
  const long long total = (long long)b * n;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
  three_nn_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(total, dist2, weight);
  count_launch();
  return launch_status();
}

MVP_API int mvp_fscore(int b, int n, int m, const float *dist1, const float *dist2, float threshold, float *fscore,
                       float *precision1, float *precision2, mvp_stream_t stream) {
  if (b < 0 || n <= 0 || m <= 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0) return MVP_OK;
  if (!dist1 || !dist2 || !fscore || !precision1 || !precision2) return MVP_ERR_INVALID_ARGUMENT;
  fscore_kernel<<<b, kLossThreads, 0, (cudaStream_t)stream>>>(n, m, threshold, dist1, dist2, fscore, precision1, precision2);
  count_launch();
  return launch_status();
}

MVP_API int mvp_chamfer_loss(int b, int n, int m, const float *dist1, const float *dist2, float *cd_p, float *cd_t,
                             mvp_stream_t stream) {
  if (b < 0 || n <= 0 || m <= 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0) return MVP_OK;
  if (!dist1 || !dist2 || !cd_p || !cd_t) return MVP_ERR_INVALID_ARGUMENT;
  chamfer_loss_kernel<<<b, kLossThreads, 0, (cudaStream_t)stream>>>(n, m, dist1, dist2, cd_p, cd_t);
  count_launch();
  return launch_status();
}

MVP_API int mvp_chamfer_loss_grad(int b, int n, int m, const float *dist1, const float *dist2, const float *grad_cd_p,
                                  const float *grad_cd_t, float *grad_dist1, float *grad_dist2, mvp_stream_t stream) {
  if (b < 0 || n <= 0 || m <= 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0) return MVP_OK;
  if (!dist1 || !dist2 || !grad_cd_p || !grad_cd_t || !grad_dist1 || !grad_dist2) return MVP_ERR_INVALID_ARGUMENT;
  const long long total = (long long)b * ((long long)n + m);
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
  chamfer_loss_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(b, n, m, dist1, dist2, grad_cd_p, grad_cd_t,
                                                                   grad_dist1, grad_dist2);
  count_launch();
  return launch_status();
}
