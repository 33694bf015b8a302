// PointNet++ neighbourhood / gather ops for sm_100a: ball_query, three_nn, knn, gather_points,
// group_points, three_interpolate (+ their backward scatters).
//
// Replaces the kernels of the reference under utils/mm3d_pn2/ops/:
//   ball_query/src/ball_query_cuda.cu:11-54        interpolate/src/three_nn_cuda.cu:11-65
//   knn/src/knn_cuda.cu:58-94                      gather_points/src/gather_points_cuda.cu:8-26,51-70
//   group_points/src/group_points_cuda.cu:10-31,56-79
//   interpolate/src/three_interpolate_cuda.cu:11-35,61-84
//
// Design differences from the reference (results are identical):
//   * searches stage the searched cloud through shared memory once per CTA instead of every thread
//     streaming it from global memory; ball_query uses a WARP per centre (32 candidates per step,
//     ballot + prefix-popcount keeps the ascending-index order and the early exit);
//   * gathers read each index (and interpolation weight) ONCE per point and loop channels inside the
//     thread — the reference re-reads them once per channel;
//   * outputs that the reference's Python zero-fills are zero-filled here.
#include "common.cuh"

namespace mvp {

constexpr int kTile = 1024;  // searched points per shared-memory tile (AoS, 12 KB)

__device__ __forceinline__ void load_tile(float *tile, const float *__restrict__ src, int cnt, int tid,
                                          int nthreads) {
  for (int i = tid; i < cnt * 3; i += nthreads) tile[i] = __ldg(src + i);
}

// ------------------------------------------------------------------------------------------------
// ball_query: one warp per centre.
// ------------------------------------------------------------------------------------------------
constexpr int kBqWarps = 8;

__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(int n, int m, float min_radius2, float max_radius2, int nsample,
                  const float *__restrict__ new_xyz, const float *__restrict__ xyz, int *__restrict__ idx,
                  float *__restrict__ grouped) {
  __shared__ float tile[kTile * 3];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * kBqWarps + warp;
  const bool active = p < m;
  const float *pts = xyz + (size_t)b * n * 3;
  float cx = 0, cy = 0, cz = 0;
  int *out = nullptr;
  if (active) {
    const float *c = new_xyz + ((size_t)b * m + p) * 3;
    cx = __ldg(c + 0);
    cy = __ldg(c + 1);
    cz = __ldg(c + 2);
    out = idx + ((size_t)b * m + p) * nsample;
  }
  int cnt = 0, first = 0;
  bool done = !active;
  for (int k2 = 0; k2 < n; k2 += kTile) {
    const int tcnt = min(kTile, n - k2);
    __syncthreads();
    load_tile(tile, pts + (size_t)k2 * 3, tcnt, threadIdx.x, kBqWarps * 32);
    __syncthreads();
    if (done) continue;
    for (int k = 0; k < tcnt && !done; k += 32) {
      const int kk = k + lane;
      bool hit = false;
      if (kk < tcnt) {
        const float d2 = sqdist(cx - tile[kk * 3 + 0], cy - tile[kk * 3 + 1], cz - tile[kk * 3 + 2]);
        hit = (d2 == 0.f) || (d2 >= min_radius2 && d2 < max_radius2);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (mask) {
        if (cnt == 0) first = k2 + k + __ffs(mask) - 1;
        const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
        if (hit && pos < nsample) out[pos] = k2 + kk;
        cnt += __popc(mask);
        if (cnt >= nsample) done = true;
      }
    }
  }
  if (active) {
    // Rows are zero when nothing was hit (ball_query.py:35); otherwise padded with the first hit (:44-48).
    const int filled = min(cnt, nsample);
    for (int l = filled + lane; l < nsample; l += 32) out[l] = first;
    if (grouped != nullptr) {
      // ball_query -> grouping_operation on the coordinates -> permute(0, 2, 3, 1) (completion/model_utils.py:211-214)
      // in the same launch: the neighbours' coordinates as (b, m, nsample, 3); a row without a hit groups point 0
      __syncwarp();
      float *g = grouped + ((size_t)b * m + p) * nsample * 3;
      for (int l = lane; l < nsample; l += 32) {
        const int k = cnt ? out[l] : 0;
        g[l * 3 + 0] = __ldg(pts + (size_t)k * 3 + 0);
        g[l * 3 + 1] = __ldg(pts + (size_t)k * 3 + 1);
        g[l * 3 + 2] = __ldg(pts + (size_t)k * 3 + 2);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// three_nn: one thread per target, sources broadcast from shared memory.
// ------------------------------------------------------------------------------------------------
// kRest: the targets are the count[b] entries of list + b*n (the points the grid search of mvp_three_nn_ws did not
// finish, chamfer_grid.cu); CTAs beyond the list length leave at once.
template <bool kRest>
__global__ void __launch_bounds__(256)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known,
                float *__restrict__ dist2, int *__restrict__ idx, const int *__restrict__ list,
                const int *__restrict__ count, float *__restrict__ weight) {
  __shared__ float tile[kTile * 3];
  const int b = blockIdx.y;
  const int nq = kRest ? __ldg(count + b) : n;
  if (blockIdx.x * 256 >= nq) return;
  int p = blockIdx.x * 256 + threadIdx.x;
  const bool active = p < nq;
  if (kRest && active) p = __ldg(list + (size_t)b * n + p);
  const float *kn = known + (size_t)b * m * 3;
  float ux = 0, uy = 0, uz = 0;
  if (active) {
    const float *u = unknown + ((size_t)b * n + p) * 3;
    ux = __ldg(u + 0);
    uy = __ldg(u + 1);
    uz = __ldg(u + 2);
  }
  // The reference compares the fp32 distance against DOUBLE bests initialised to 1e40 and narrows on
  // store (three_nn_cuda.cu:35,59-61); +inf in fp32 gives the same decisions and the same stored value.
  const float inf = __int_as_float(0x7f800000);
  float best1 = inf, best2 = inf, best3 = inf;
  int besti1 = 0, besti2 = 0, besti3 = 0;
  for (int k2 = 0; k2 < m; k2 += kTile) {
    const int tcnt = min(kTile, m - k2);
    __syncthreads();
    load_tile(tile, kn + (size_t)k2 * 3, tcnt, threadIdx.x, 256);
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < tcnt; ++k) {
      const float d = sqdist(ux - tile[k * 3 + 0], uy - tile[k * 3 + 1], uz - tile[k * 3 + 2]);
      if (d < best3) {
        const int kk = k2 + k;
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = kk;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = kk;
        } else {
          best3 = d; besti3 = kk;
        }
      }
    }
  }
  if (active) {
    float *od = dist2 + ((size_t)b * n + p) * 3;
    int *oi = idx + ((size_t)b * n + p) * 3;
    od[0] = best1; od[1] = best2; od[2] = best3;
    oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
    if (weight != nullptr) three_nn_weights_of(best1, best2, best3, weight + ((size_t)b * n + p) * 3);
  }
}

// ------------------------------------------------------------------------------------------------
// knn (replaces knn_cuda.cu:58-94: one thread per centre, a 100-slot max-heap in local memory, heap sort).
// A WARP per centre, LANES OVER CANDIDATES: the searched cloud goes through a shared-memory tile (structure of
// arrays) shared by the eight centres of a CTA; a lane evaluates one candidate per step, and the (up to) 32 best so far
// are an ORDERED LIST HELD ACROSS THE LANES — lane i keeps the i-th nearest (distance, index).  A step first filters
// its 32 candidates against the current k-th entry (one ballot: nearly always empty once the list has warmed up); a
// survivor is inserted with one ballot (its rank = how many entries precede it) and one shuffle-up of the tail.
// No local memory, no heap sort at the end, m x 32 threads instead of m.  nsample > 32 (the reference allows 100) runs
// further passes, each keeping only candidates AFTER the last entry of the previous pass in (distance, index) order.
// Same result as the reference: the k smallest by (distance, index) — its strict `d2 < best_dist[0]` keeps the
// earlier index among equal distances at the boundary (knn_cuda.cu:83) — in ascending distance; entries of EQUAL
// distance come out in index order here and in heap order there.  Slots the cloud cannot fill
// (n < nsample) keep the reference's initial (1e10, 0) (knn_cuda.cu:74-77).  NaN distances are never kept, as there.
// ------------------------------------------------------------------------------------------------
constexpr int kKnnWarps = 8;

// output slots [done, done + kk) of every centre, kk <= 32; done > 0: only candidates after slot done - 1
__global__ void __launch_bounds__(kKnnWarps * 32)
knn_warp_kernel(int n, int m, int nsample, int done, int kk, const float *__restrict__ xyz,
                const float *__restrict__ new_xyz, int *__restrict__ idx, float *__restrict__ dist2) {
  __shared__ float tx[kTile], ty[kTile], tz[kTile];
  const unsigned full = 0xffffffffu;
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * kKnnWarps + warp;
  const bool active = p < m;
  const float *pts = xyz + (size_t)b * n * 3;
  const float inf = __int_as_float(0x7f800000);
  float cx = 0, cy = 0, cz = 0, ld0 = -inf;
  int li0 = -1;
  int *oi = idx + ((size_t)b * m + (active ? p : 0)) * nsample;
  float *od = dist2 + ((size_t)b * m + (active ? p : 0)) * nsample;
  if (active) {
    const float *c = new_xyz + ((size_t)b * m + p) * 3;
    cx = __ldg(c + 0), cy = __ldg(c + 1), cz = __ldg(c + 2);
    if (done) ld0 = od[done - 1], li0 = oi[done - 1];
  }
  float ld = 1e10f, thr_d = 1e10f;  // this lane's entry of the list; the kk-th entry
  int li = 0, thr_i = 0;
  for (int j0 = 0; j0 < n; j0 += kTile) {
    const int cnt = min(kTile, n - j0);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += kKnnWarps * 32) {
      const float *c = pts + (size_t)(j0 + j) * 3;
      tx[j] = __ldg(c + 0), ty[j] = __ldg(c + 1), tz[j] = __ldg(c + 2);
    }
    __syncthreads();
    if (!active) continue;
    for (int t0 = 0; t0 < cnt; t0 += 32) {
      const int t = t0 + lane, j = j0 + t;
      bool pass = false;
      float d = 0.f;
      if (t < cnt) {
        d = sqdist(cx - tx[t], cy - ty[t], cz - tz[t]);
        pass = (d < thr_d || (d == thr_d && j < thr_i)) && (d > ld0 || (d == ld0 && j > li0));
      }
      unsigned mask = __ballot_sync(full, pass);
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const float cd = __shfl_sync(full, d, src);
        const int cj = j0 + t0 + src;
        if (!(cd < thr_d || (cd == thr_d && cj < thr_i))) continue;  // the k-th entry has moved since the filter
        const int pos = __popc(__ballot_sync(full, ld < cd || (ld == cd && li < cj)));
        const float ud = __shfl_up_sync(full, ld, 1);
        const int ui = __shfl_up_sync(full, li, 1);
        if (lane > pos) ld = ud, li = ui;
        else if (lane == pos) ld = cd, li = cj;
        thr_d = __shfl_sync(full, ld, kk - 1);
        thr_i = __shfl_sync(full, li, kk - 1);
      }
    }
  }
  if (active && lane < kk) {
    od[done + lane] = ld;
    oi[done + lane] = li;
  }
}

// ------------------------------------------------------------------------------------------------
// top-k of the rows of a score matrix (the feature-space case): completion/model_utils.py:242-247
// ranks the neighbours of a dynamic graph by a (B, N, N) matrix  -|x|^2 + 2 x.y - |y|^2  on C-dimensional FEATURES and
// calls torch.topk, whose multi-block radix select is 6.4 ms of a 35 ms ECG step.  A warp per row: lanes stream the
// row coalesced, the k <= 32 LARGEST so far are an ordered list held across the lanes exactly as in knn_warp_kernel
// (one ballot to filter a step's 32 scores against the k-th, one ballot + one shuffle-up per insertion).  Output:
// values and indices in DESCENDING value, equal values in ascending index (torch.topk leaves that order
// unspecified); NaN scores are never selected.
// ------------------------------------------------------------------------------------------------
// kScore: `scores` is the Gram matrix x^T x of a cloud's feature vectors ((B, N, N), rows = B * N, cols = N) and the
// ranked score is the original's  -|x_j|^2 - (-2 x_i.x_j) - |x_i|^2  (completion/model_utils.py:243-245), each step the
// IEEE operation torch performs in its three elementwise kernels (x -2 is exact; two subtractions in that order): the
// (B, N, N) score matrix is never written or re-read.
constexpr int kTopkWarps = 8;
template <bool kScore>
__global__ void __launch_bounds__(kTopkWarps * 32)
topk_rows_kernel(long long rows, int cols, int k, const float *__restrict__ scores, const float *__restrict__ sqnorm,
                 float *__restrict__ vals, long long *__restrict__ idx64, int *__restrict__ idx32) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)kTopkWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float *r = scores + row * cols;
  const float *nb = nullptr;  // the cloud's squared norms
  float own = 0.f;
  if (kScore) {
    nb = sqnorm + (row / cols) * cols;
    own = __ldg(nb + (row % cols));
  }
  const float ninf = __int_as_float(0xff800000);
  float lv = ninf, thr_v = ninf;  // this lane's entry; the k-th entry
  int li = 0x7fffffff, thr_i = 0x7fffffff;
  // four steps of 32 scores in flight together (the row streams from HBM: a dependent 4-byte load per step would leave
  // the memory system idle); each step's 32 scores are then filtered and inserted in column order
  for (int t0 = 0; t0 < cols; t0 += 128) {
    float v4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int j = t0 + 32 * u + lane;
      v4[u] = j < cols ? __ldg(r + j) : ninf;
      if (kScore && j < cols) v4[u] = __fsub_rn(__fsub_rn(-__ldg(nb + j), __fmul_rn(-2.f, v4[u])), own);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int j = t0 + 32 * u + lane;
      const float v = v4[u];
      const bool pass = j < cols && (v > thr_v || (v == thr_v && j < thr_i));
      unsigned mask = __ballot_sync(full, pass);
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const float cv = __shfl_sync(full, v, src);
        const int cj = t0 + 32 * u + src;
        if (!(cv > thr_v || (cv == thr_v && cj < thr_i))) continue;  // the k-th entry has moved since the filter
        const int pos = __popc(__ballot_sync(full, lv > cv || (lv == cv && li < cj)));
        const float uv = __shfl_up_sync(full, lv, 1);
        const int ui = __shfl_up_sync(full, li, 1);
        if (lane > pos) lv = uv, li = ui;
        else if (lane == pos) lv = cv, li = cj;
        thr_v = __shfl_sync(full, lv, k - 1);
        thr_i = __shfl_sync(full, li, k - 1);
      }
    }
  }
  if (lane < k) {
    const int out = li == 0x7fffffff ? lane : li;  // fewer than k comparable scores (NaN / -inf rows): a valid index
    if (vals) vals[row * k + lane] = lv;
    if (idx64) idx64[row * k + lane] = out;
    if (idx32) idx32[row * k + lane] = out;
  }
}

// ------------------------------------------------------------------------------------------------
// gather_points / group_points (group is gather with npoints*nsample indices per cloud).
// grid (ceil(M/256), csplit, B): a thread owns one output column p and walks a slice of channels;
// stores along p are coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_kernel(int c, int n, int mpts, int cper, const float *__restrict__ points,
              const int *__restrict__ idx, float *__restrict__ out) {
  const int b = blockIdx.z;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= mpts) return;
  const int c0 = blockIdx.y * cper, c1 = min(c, c0 + cper);
  const int src = __ldg(idx + (size_t)b * mpts + p);
  const float *pp = points + ((size_t)b * c + c0) * n + src;
  float *oo = out + ((size_t)b * c + c0) * mpts + p;
#pragma unroll 4
  for (int ci = c0; ci < c1; ci++) {
    *oo = __ldg(pp);
    pp += n;
    oo += mpts;
  }
}

__global__ void __launch_bounds__(256)
gather_grad_kernel(int c, int n, int mpts, int cper, const float *__restrict__ grad_out,
                   const int *__restrict__ idx, float *__restrict__ grad_points) {
  const int b = blockIdx.z;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= mpts) return;
  const int c0 = blockIdx.y * cper, c1 = min(c, c0 + cper);
  const int dst = __ldg(idx + (size_t)b * mpts + p);
  float *gp = grad_points + ((size_t)b * c + c0) * n + dst;
  const float *go = grad_out + ((size_t)b * c + c0) * mpts + p;
#pragma unroll 4
  for (int ci = c0; ci < c1; ci++) {
    atomicAdd(gp, __ldg(go));
    gp += n;
    go += mpts;
  }
}

// ------------------------------------------------------------------------------------------------
// three_interpolate: out = fma(w2,p2, fma(w0,p0, w1*p1)) — the contraction nvcc gives
// three_interpolate_cuda.cu:33-34 (SASS-verified).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
three_interpolate_kernel(int c, int m, int n, int cper, const float *__restrict__ points,
                         const int *__restrict__ idx, const float *__restrict__ weight,
                         float *__restrict__ out) {
  const int b = blockIdx.z;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= n) return;
  const int c0 = blockIdx.y * cper, c1 = min(c, c0 + cper);
  const int *id = idx + ((size_t)b * n + p) * 3;
  const float *w = weight + ((size_t)b * n + p) * 3;
  const int i0 = __ldg(id + 0), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
  const float w0 = __ldg(w + 0), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
  const float *pp = points + ((size_t)b * c + c0) * m;
  float *oo = out + ((size_t)b * c + c0) * n + p;
#pragma unroll 4
  for (int ci = c0; ci < c1; ci++) {
    *oo = __fmaf_rn(w2, __ldg(pp + i2), __fmaf_rn(w0, __ldg(pp + i0), __fmul_rn(w1, __ldg(pp + i1))));
    pp += m;
    oo += n;
  }
}

__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(int c, int n, int m, int cper, const float *__restrict__ grad_out,
                              const int *__restrict__ idx, const float *__restrict__ weight,
                              float *__restrict__ grad_points) {
  const int b = blockIdx.z;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= n) return;
  const int c0 = blockIdx.y * cper, c1 = min(c, c0 + cper);
  const int *id = idx + ((size_t)b * n + p) * 3;
  const float *w = weight + ((size_t)b * n + p) * 3;
  const int i0 = __ldg(id + 0), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
  const float w0 = __ldg(w + 0), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
  float *gp = grad_points + ((size_t)b * c + c0) * m;
  const float *go = grad_out + ((size_t)b * c + c0) * n + p;
#pragma unroll 4
  for (int ci = c0; ci < c1; ci++) {
    const float g = __ldg(go);
    atomicAdd(gp + i0, __fmul_rn(g, w0));
    atomicAdd(gp + i1, __fmul_rn(g, w1));
    atomicAdd(gp + i2, __fmul_rn(g, w2));
    gp += m;
    go += n;
  }
}

// Channel slices: enough CTAs to fill the machine (>= 4 per SM) without making slices tiny.
static int channel_slices(int b, int c, int cols) {
  const long long col_ctas = (long long)b * ((cols + 255) / 256);
  int split = 1;
  while (split < c && col_ctas * split < 4LL * kNumSMs && c / (split * 2) >= 4) split *= 2;
  return split;
}

static bool bad_dims(int b, int c, int n, int mpts) { return b < 0 || c < 0 || n < 0 || mpts < 0; }

// pointnet2_staged.cu: the shared-memory staged variants, taken when there are about as many gathered columns as
// source columns (every call of the completion models)
bool staged_applicable(int b, int c, int rows, int cols);
int gather_staged_launch(int b, int c, int n, int mpts, const float *points, const int *idx, float *out,
                         cudaStream_t s);
int gather_grad_staged_launch(int b, int c, int n, int mpts, const float *grad_out, const int *idx, float *grad_points,
                              cudaStream_t s);
int three_interpolate_staged_launch(int b, int c, int m, int n, const float *points, const int *idx, const float *weight,
                                    float *out, cudaStream_t s);
int three_interpolate_grad_staged_launch(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                         const float *weight, float *grad_points, cudaStream_t s);
size_t scatter_csr_workspace_bytes(int b, int rows, int entries);
int scatter_csr_launch(bool interp, int b, int c, int rows, int cols, const float *grad_out, const int *idx,
                       const float *weight, float *grad_points, void *workspace, size_t workspace_bytes, cudaStream_t s);

static int gather_launch(int b, int c, int n, int mpts, const float *points, const int *idx, float *out,
                         cudaStream_t s) {
  if (bad_dims(b, c, n, mpts)) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || c == 0 || mpts == 0) return MVP_OK;
  if (n == 0 || !points || !idx || !out) return MVP_ERR_INVALID_ARGUMENT;
  if (staged_applicable(b, c, n, mpts)) return gather_staged_launch(b, c, n, mpts, points, idx, out, s);
  const int split = channel_slices(b, c, mpts);
  const int cper = (c + split - 1) / split;
  for (int b0 = 0; b0 < b; b0 += 65535) {
    const int bb = min(65535, b - b0);
    dim3 grid((mpts + 255) / 256, (c + cper - 1) / cper, bb);
    gather_kernel<<<grid, 256, 0, s>>>(c, n, mpts, cper, points + (size_t)b0 * c * n,
                                       idx + (size_t)b0 * mpts, out + (size_t)b0 * c * mpts);
    count_launch();
  }
  return launch_status();
}

static int gather_grad_launch(int b, int c, int n, int mpts, const float *grad_out, const int *idx,
                              float *grad_points, cudaStream_t s) {
  if (bad_dims(b, c, n, mpts)) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || c == 0 || n == 0) return MVP_OK;
  if (!grad_points) return MVP_ERR_INVALID_ARGUMENT;
  if (mpts > 0 && grad_out && idx && staged_applicable(b, c, n, mpts))  // writes every element: no memset
    return gather_grad_staged_launch(b, c, n, mpts, grad_out, idx, grad_points, s);
  cudaError_t e = cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * n, s);
  if (e != cudaSuccess) return (int)e;
  if (mpts == 0) return MVP_OK;
  if (!grad_out || !idx) return MVP_ERR_INVALID_ARGUMENT;
  const int split = channel_slices(b, c, mpts);
  const int cper = (c + split - 1) / split;
  for (int b0 = 0; b0 < b; b0 += 65535) {
    const int bb = min(65535, b - b0);
    dim3 grid((mpts + 255) / 256, (c + cper - 1) / cper, bb);
    gather_grad_kernel<<<grid, 256, 0, s>>>(c, n, mpts, cper, grad_out + (size_t)b0 * c * mpts,
                                            idx + (size_t)b0 * mpts, grad_points + (size_t)b0 * c * n);
    count_launch();
  }
  return launch_status();
}

}  // namespace mvp

using namespace mvp;

static int ball_query_launch(int b, int n, int m, float min_radius, float max_radius, int nsample, const float *new_xyz,
                             const float *xyz, int *idx, float *grouped, mvp_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || nsample < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || m == 0 || nsample == 0) return MVP_OK;
  if (!new_xyz || !idx || (n > 0 && !xyz)) return MVP_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(idx, 0, sizeof(int) * (size_t)b * m * nsample, s);  // ball_query.py:35
  if (e != cudaSuccess) return (int)e;
  if (n == 0) return grouped ? MVP_ERR_INVALID_ARGUMENT : MVP_OK;  // nothing to group from
  // Radii are squared in fp32 exactly as ball_query_cuda.cu:30-31 does.
  const float max_radius2 = max_radius * max_radius;
  const float min_radius2 = min_radius * min_radius;
  for (int b0 = 0; b0 < b; b0 += 65535) {
    const int bb = min(65535, b - b0);
    dim3 grid((m + kBqWarps - 1) / kBqWarps, bb);
    ball_query_kernel<<<grid, kBqWarps * 32, 0, s>>>(n, m, min_radius2, max_radius2, nsample,
                                                     new_xyz + (size_t)b0 * m * 3, xyz + (size_t)b0 * n * 3,
                                                     idx + (size_t)b0 * m * nsample,
                                                     grouped ? grouped + (size_t)b0 * m * nsample * 3 : nullptr);
    count_launch();
  }
  return launch_status();
}

MVP_API int mvp_ball_query(int b, int n, int m, float min_radius, float max_radius, int nsample,
                           const float *new_xyz, const float *xyz, int *idx, mvp_stream_t stream) {
  return ball_query_launch(b, n, m, min_radius, max_radius, nsample, new_xyz, xyz, idx, nullptr, stream);
}

// ball_query + the neighbours' coordinates (b, m, nsample, 3) in one launch
MVP_API int mvp_ball_query_group(int b, int n, int m, float min_radius, float max_radius, int nsample,
                                 const float *new_xyz, const float *xyz, int *idx, float *grouped_xyz,
                                 mvp_stream_t stream) {
  if (!grouped_xyz && b > 0 && m > 0 && nsample > 0) return MVP_ERR_INVALID_ARGUMENT;
  return ball_query_launch(b, n, m, min_radius, max_radius, nsample, new_xyz, xyz, idx, grouped_xyz, stream);
}

static int three_nn_exhaustive(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                               float *weight, mvp_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0) return MVP_OK;
  if (!unknown || !dist2 || !idx || (m > 0 && !known)) return MVP_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  for (int b0 = 0; b0 < b; b0 += 65535) {
    const int bb = min(65535, b - b0);
    dim3 grid((n + 255) / 256, bb);
    three_nn_kernel<false><<<grid, 256, 0, s>>>(n, m, unknown + (size_t)b0 * n * 3, known + (size_t)b0 * m * 3,
                                                dist2 + (size_t)b0 * n * 3, idx + (size_t)b0 * n * 3, nullptr, nullptr,
                                                weight ? weight + (size_t)b0 * n * 3 : nullptr);
    count_launch();
  }
  return launch_status();
}

MVP_API int mvp_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                         int *idx, mvp_stream_t stream) {
  return three_nn_exhaustive(b, n, m, unknown, known, dist2, idx, nullptr, stream);
}

namespace mvp {
// chamfer_grid.cu
bool three_nn_grid_supported(int b, int n, int m);
size_t three_nn_grid_workspace_bytes(int b, int n, int m);
int three_nn_grid_launch(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                         float *weight, void *ws, size_t ws_bytes, cudaStream_t s);

int three_nn_rest_launch(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                         float *weight, const int *list, const int *count, cudaStream_t s) {
  three_nn_kernel<true><<<dim3((n + 255) / 256, b), 256, 0, s>>>(n, m, unknown, known, dist2, idx, list, count, weight);
  count_launch();
  return launch_status();
}
}  // namespace mvp

MVP_API size_t mvp_three_nn_workspace_bytes(int b, int n, int m) {
  return three_nn_grid_supported(b, n, m) ? three_nn_grid_workspace_bytes(b, n, m) : 16;
}

static int three_nn_ws(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                       float *weight, void *workspace, size_t workspace_bytes, mvp_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (three_nn_grid_supported(b, n, m) && unknown && known && dist2 && idx && workspace &&
      workspace_bytes >= three_nn_grid_workspace_bytes(b, n, m))
    return three_nn_grid_launch(b, n, m, unknown, known, dist2, idx, weight, workspace, workspace_bytes,
                                (cudaStream_t)stream);
  return three_nn_exhaustive(b, n, m, unknown, known, dist2, idx, weight, stream);
}

MVP_API int mvp_three_nn_ws(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                            void *workspace, size_t workspace_bytes, mvp_stream_t stream) {
  return three_nn_ws(b, n, m, unknown, known, dist2, idx, nullptr, workspace, workspace_bytes, stream);
}

// three_nn and the inverse-distance weights of three_nn_upsampling (completion/model_utils.py:286-293) from the same
// launches: the kernel that finds a target's three neighbours also writes its weights.
MVP_API int mvp_three_nn_weights_ws(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                                    float *weight, void *workspace, size_t workspace_bytes, mvp_stream_t stream) {
  if (!weight && b > 0 && n > 0) return MVP_ERR_INVALID_ARGUMENT;
  return three_nn_ws(b, n, m, unknown, known, dist2, idx, weight, workspace, workspace_bytes, stream);
}

MVP_API int mvp_knn(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx,
                    float *dist2, mvp_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (nsample <= 0 || nsample > 100) return MVP_ERR_INVALID_ARGUMENT;  // knn_cuda.cu:72-73 (100 slots)
  if (b == 0 || m == 0) return MVP_OK;
  if (!new_xyz || !idx || !dist2 || (n > 0 && !xyz)) return MVP_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  for (int done = 0; done < nsample; done += 32) {
    const int kk = min(32, nsample - done);
    for (int b0 = 0; b0 < b; b0 += 65535) {
      const int bb = min(65535, b - b0);
      knn_warp_kernel<<<dim3((m + kKnnWarps - 1) / kKnnWarps, bb), kKnnWarps * 32, 0, s>>>(
          n, m, nsample, done, kk, xyz + (size_t)b0 * n * 3, new_xyz + (size_t)b0 * m * 3, idx + (size_t)b0 * m * nsample,
          dist2 + (size_t)b0 * m * nsample);
      count_launch();
    }
  }
  return launch_status();
}

MVP_API int mvp_topk_rows(long long rows, int cols, int k, const float *scores, float *values, long long *idx64,
                          int *idx32, mvp_stream_t stream) {
  if (rows < 0 || cols <= 0 || k <= 0 || k > 32 || k > cols) return MVP_ERR_INVALID_ARGUMENT;
  if (rows == 0) return MVP_OK;
  if (!scores || (!values && !idx64 && !idx32)) return MVP_ERR_INVALID_ARGUMENT;
  const long long blocks = (rows + kTopkWarps - 1) / kTopkWarps;
  if (blocks > 0x7fffffffLL) return MVP_ERR_INVALID_ARGUMENT;
  topk_rows_kernel<false><<<(unsigned)blocks, kTopkWarps * 32, 0, (cudaStream_t)stream>>>(rows, cols, k, scores, nullptr,
                                                                                           values, idx64, idx32);
  count_launch();
  return launch_status();
}

MVP_API int mvp_topk_rows_sqdist(int b, int n, int k, const float *gram, const float *sqnorm, float *values,
                                 long long *idx64, int *idx32, mvp_stream_t stream) {
  if (b < 0 || n <= 0 || k <= 0 || k > 32 || k > n) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0) return MVP_OK;
  if (!gram || !sqnorm || (!values && !idx64 && !idx32)) return MVP_ERR_INVALID_ARGUMENT;
  const long long rows = (long long)b * n, blocks = (rows + kTopkWarps - 1) / kTopkWarps;
  if (blocks > 0x7fffffffLL) return MVP_ERR_INVALID_ARGUMENT;
  topk_rows_kernel<true><<<(unsigned)blocks, kTopkWarps * 32, 0, (cudaStream_t)stream>>>(rows, n, k, gram, sqnorm, values,
                                                                                          idx64, idx32);
  count_launch();
  return launch_status();
}

MVP_API int mvp_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx,
                              float *out, mvp_stream_t stream) {
  return gather_launch(b, c, n, npoints, points, idx, out, (cudaStream_t)stream);
}

MVP_API int mvp_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                                   float *grad_points, mvp_stream_t stream) {
  return gather_grad_launch(b, c, n, npoints, grad_out, idx, grad_points, (cudaStream_t)stream);
}

MVP_API size_t mvp_scatter_workspace_bytes(int b, int rows, int entries) {
  return scatter_csr_workspace_bytes(b, rows, entries);
}

// With a workspace the backward scatters go through a transposed index (pointnet2_staged.cu); shapes that path does
// not cover fall back to the workspace-free kernels.
static int gather_grad_ws(int b, int c, int n, int mpts, const float *grad_out, const int *idx, float *grad_points,
                          void *workspace, size_t workspace_bytes, cudaStream_t s) {
  if (bad_dims(b, c, n, mpts)) return MVP_ERR_INVALID_ARGUMENT;
  if (b > 0 && c > 0 && n > 0 && mpts > 0 && grad_out && idx && grad_points) {
    const int rc = scatter_csr_launch(false, b, c, n, mpts, grad_out, idx, nullptr, grad_points, workspace,
                                      workspace_bytes, s);
    if (rc != -100) return rc;
  }
  return gather_grad_launch(b, c, n, mpts, grad_out, idx, grad_points, s);
}

MVP_API int mvp_gather_points_grad_ws(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                                      float *grad_points, void *workspace, size_t workspace_bytes, mvp_stream_t stream) {
  return gather_grad_ws(b, c, n, npoints, grad_out, idx, grad_points, workspace, workspace_bytes, (cudaStream_t)stream);
}

MVP_API int mvp_group_points_grad_ws(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                     const int *idx, float *grad_points, void *workspace, size_t workspace_bytes,
                                     mvp_stream_t stream) {
  if (npoints < 0 || nsample < 0 || (long long)npoints * nsample > 0x7fffffffLL)
    return MVP_ERR_INVALID_ARGUMENT;
  return gather_grad_ws(b, c, n, npoints * nsample, grad_out, idx, grad_points, workspace, workspace_bytes,
                        (cudaStream_t)stream);
}

MVP_API int mvp_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                             const int *idx, float *out, mvp_stream_t stream) {
  if (npoints < 0 || nsample < 0 || (long long)npoints * nsample > 0x7fffffffLL)
    return MVP_ERR_INVALID_ARGUMENT;
  return gather_launch(b, c, n, npoints * nsample, points, idx, out, (cudaStream_t)stream);
}

MVP_API int mvp_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                  const int *idx, float *grad_points, mvp_stream_t stream) {
  if (npoints < 0 || nsample < 0 || (long long)npoints * nsample > 0x7fffffffLL)
    return MVP_ERR_INVALID_ARGUMENT;
  return gather_grad_launch(b, c, n, npoints * nsample, grad_out, idx, grad_points, (cudaStream_t)stream);
}

MVP_API int mvp_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                  const float *weight, float *out, mvp_stream_t stream) {
  if (bad_dims(b, c, m, n)) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || c == 0 || n == 0) return MVP_OK;
  if (m == 0 || !points || !idx || !weight || !out) return MVP_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  if (staged_applicable(b, c, m, n)) return three_interpolate_staged_launch(b, c, m, n, points, idx, weight, out, s);
  const int split = channel_slices(b, c, n);
  const int cper = (c + split - 1) / split;
  for (int b0 = 0; b0 < b; b0 += 65535) {
    const int bb = min(65535, b - b0);
    dim3 grid((n + 255) / 256, (c + cper - 1) / cper, bb);
    three_interpolate_kernel<<<grid, 256, 0, s>>>(c, m, n, cper, points + (size_t)b0 * c * m,
                                                  idx + (size_t)b0 * n * 3, weight + (size_t)b0 * n * 3,
                                                  out + (size_t)b0 * c * n);
    count_launch();
  }
  return launch_status();
}

MVP_API int mvp_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                       const float *weight, float *grad_points, mvp_stream_t stream);

MVP_API int mvp_three_interpolate_grad_ws(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                          const float *weight, float *grad_points, void *workspace,
                                          size_t workspace_bytes, mvp_stream_t stream) {
  if (bad_dims(b, c, m, n)) return MVP_ERR_INVALID_ARGUMENT;
  if (b > 0 && c > 0 && m > 0 && n > 0 && (long long)n * 3 <= 0x7fffffffLL && grad_out && idx && weight && grad_points) {
    const int rc = scatter_csr_launch(true, b, c, m, n, grad_out, idx, weight, grad_points, workspace, workspace_bytes,
                                      (cudaStream_t)stream);
    if (rc != -100) return rc;
  }
  return mvp_three_interpolate_grad(b, c, n, m, grad_out, idx, weight, grad_points, stream);
}

MVP_API int mvp_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                       const float *weight, float *grad_points, mvp_stream_t stream) {
  if (bad_dims(b, c, m, n)) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || c == 0 || m == 0) return MVP_OK;
  if (!grad_points) return MVP_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  if (n > 0 && grad_out && idx && weight && staged_applicable(b, c, m, n))  // writes every element: no memset
    return three_interpolate_grad_staged_launch(b, c, n, m, grad_out, idx, weight, grad_points, s);
  cudaError_t e = cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * m, s);
  if (e != cudaSuccess) return (int)e;
  if (n == 0) return MVP_OK;
  if (!grad_out || !idx || !weight) return MVP_ERR_INVALID_ARGUMENT;
  const int split = channel_slices(b, c, n);
  const int cper = (c + split - 1) / split;
  for (int b0 = 0; b0 < b; b0 += 65535) {
    const int bb = min(65535, b - b0);
    dim3 grid((n + 255) / 256, (c + cper - 1) / cper, bb);
    three_interpolate_grad_kernel<<<grid, 256, 0, s>>>(c, n, m, cper, grad_out + (size_t)b0 * c * n,
                                                       idx + (size_t)b0 * n * 3, weight + (size_t)b0 * n * 3,
                                                       grad_points + (size_t)b0 * c * m);
    count_launch();
  }
  return launch_status();
}

// ---- the maximum over a point's k neighbours: the last axis of a contiguous (rows, k) view ---------------------------------
// `y, _ = torch.max(y, 3)` on the (B, C, N, k) neighbour tensors of completion/models/ecg.py:64 and
// completion/model_utils.py:53,104.  torch reduces the 16-wide inner axis with its generic reduce kernel at 0.36 TB/s
// (604 MB in 1.67 ms); here a thread owns a row (k <= 64: 16 floats = four 16-byte loads), keeps the first arg-max as
// torch does (NaN wins, first NaN), and the backward writes the row back whole — no memset, no scatter.
namespace mvp {
__global__ void __launch_bounds__(256) max_last_kernel(long long rows, int k, const float *__restrict__ x,
                                                       float *__restrict__ out, unsigned char *__restrict__ arg) {
  const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  const float *xr = x + r * k;
  float best = 0.f;
  int bi = 0;
  auto take = [&](float v, int j) {
    if (j == 0 || v > best || (v != v && best == best)) best = v, bi = j;
  };
  if ((k & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const float4 *x4 = reinterpret_cast<const float4 *>(xr);
    for (int j = 0; j < (k >> 2); j++) {
      const float4 v = __ldg(x4 + j);
      take(v.x, 4 * j), take(v.y, 4 * j + 1), take(v.z, 4 * j + 2), take(v.w, 4 * j + 3);
    }
  } else {
    for (int j = 0; j < k; j++) take(__ldg(xr + j), j);
  }
  out[r] = best;
  arg[r] = (unsigned char)bi;
}
__global__ void __launch_bounds__(256) max_last_grad_kernel(long long rows, int k, const float *__restrict__ g,
                                                            const unsigned char *__restrict__ arg, float *__restrict__ gx) {
  const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  const float gv = __ldg(g + r);
  const int a = arg[r];
  float *gr = gx + r * k;
  if ((k & 3) == 0 && (reinterpret_cast<uintptr_t>(gx) & 15) == 0) {
    float4 *g4 = reinterpret_cast<float4 *>(gr);
    for (int j = 0; j < (k >> 2); j++)
      g4[j] = make_float4(a == 4 * j ? gv : 0.f, a == 4 * j + 1 ? gv : 0.f, a == 4 * j + 2 ? gv : 0.f, a == 4 * j + 3 ? gv : 0.f);
  } else {
    for (int j = 0; j < k; j++) gr[j] = a == j ? gv : 0.f;
  }
}
}  // namespace mvp

MVP_API int mvp_max_last(long long rows, int k, const float *x, float *out, unsigned char *arg, mvp_stream_t stream) {
  if (rows < 0 || k <= 0 || k > 255) return MVP_ERR_INVALID_ARGUMENT;
  if (rows == 0) return MVP_OK;
  if (!x || !out || !arg || rows > 256LL * 2147483647LL) return MVP_ERR_INVALID_ARGUMENT;
  mvp::max_last_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, k, x, out, arg);
  mvp::count_launch();
  return mvp::launch_status();
}

MVP_API int mvp_max_last_grad(long long rows, int k, const float *grad_out, const unsigned char *arg, float *grad_x,
                              mvp_stream_t stream) {
  if (rows < 0 || k <= 0 || k > 255) return MVP_ERR_INVALID_ARGUMENT;
  if (rows == 0) return MVP_OK;
  if (!grad_out || !grad_x || !arg || rows > 256LL * 2147483647LL) return MVP_ERR_INVALID_ARGUMENT;
  mvp::max_last_grad_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, k, grad_out, arg, grad_x);
  mvp::count_launch();
  return mvp::launch_status();
}
