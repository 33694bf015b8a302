// Furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel / furthest_point_sampling_with_dist_kernel of the reference
// (utils/mm3d_pn2/ops/furthest_point_sample/src/furthest_point_sample_cuda.cu:26-141, 214-331).
//
// The reference keeps the running minimum distances in global memory and does an 11-barrier tree per
// selected point.  Here one CTA owns one cloud, every thread keeps its P points AND their running
// minima in registers for the whole scan, and one selection step costs ONE block barrier:
//   in-thread max  ->  redux.sync.max over the warp  ->  one 8-byte slot per warp in shared memory
//   -> __syncthreads -> every warp reduces the <=32 slots redundantly (no second barrier).
//
// Bit-exact tie rule.  The reference thread t owns k = t (mod T), T = opt_n_threads(n) (:11-15), takes
// the lowest k on equal distance inside a thread (strict `>`, :69-70) and its tree keeps the LOWER
// position on equality with strides T/2..1 (:17-23), so among equal maxima the winner is the point with
// the smallest  key(k) = (bitrev_T(k mod T), k div T).  We reduce (distance, key) with max-then-min.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "grid.cuh"

#ifndef MVP_FPS_PMAX
#define MVP_FPS_PMAX 8  // tuning knob (tools/pair_variants.py): points per thread before more warps are used
#endif

#ifdef MVP_FPS_TIMING  // debugging aid: clock64 phase totals of thread 0 of cloud 0, printed at the end
#include <cstdio>
#define MVP_FPS_STAMP(k) { const long long t_ = clock64(); fps_acc[k] += t_ - fps_last; fps_last = t_; }
#else
#define MVP_FPS_STAMP(k)
#endif

namespace mvp {

__device__ __forceinline__ uint32_t fps_key(int k, int log2T) {
  const uint32_t t = (uint32_t)k & ((1u << log2T) - 1u);
  const uint32_t rev = log2T ? (__brev(t) >> (32 - log2T)) : 0u;
  return (rev << 20) | ((uint32_t)k >> log2T);
}
__device__ __forceinline__ int fps_unkey(uint32_t key, int log2T) {
  const uint32_t rev = key >> 20;
  const uint32_t t = log2T ? (__brev(rev) >> (32 - log2T)) : 0u;
  return (int)(((key & 0xfffffu) << log2T) | t);
}

__device__ __forceinline__ int redux_max_s32(int v) {
  int r;
  asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}

typedef unsigned long long fps_u64;
__device__ __forceinline__ fps_u64 fps_pack2(float lo, float hi) {
  fps_u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void fps_unpack2(fps_u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// squared distances of two points (packed) to one point (duplicated): the reference's contraction
// fma(dz,dz, fma(dx,dx, dy*dy)) lane by lane, in 6 packed fp32x2 issue slots instead of 12
__device__ __forceinline__ fps_u64 fps_dist2(fps_u64 X, fps_u64 Y, fps_u64 Z, fps_u64 qx, fps_u64 qy, fps_u64 qz) {
  fps_u64 dx, dy, dz, t;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(X), "l"(qx));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(Y), "l"(qy));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(Z), "l"(qz));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(t) : "l"(dy));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(t) : "l"(dx), "l"(t));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(t) : "l"(dz), "l"(t));
  return t;
}
__device__ __forceinline__ float fps_max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Outputs of a sampling call: the indices, and optionally the sampled POINTS themselves — the sampling idiom of the
// completion models is furthest_point_sample followed by gather_points on the transposed cloud and a transpose back
// (completion/model_utils.py:91-93, vrcnet.py:451); with `xyz` set the kernel that chose the points also writes their
// coordinates, (b, m, 3) or channels-first (b, 3, m), as an epilogue of the CTA that owns the cloud.
struct FpsOut {
  int *idx;
  float *xyz;  // nullptr: indices only
  int cf;      // 1: (b, 3, m), what gather_points returns; 0: (b, m, 3)
};
__device__ __forceinline__ void fps_write_points(const FpsOut &O, int cloud, int m, const int *idxs,
                                                 const float *__restrict__ dataset, int tid, int nthreads) {
  if (O.xyz == nullptr) return;
  __syncthreads();  // thread 0's index stores are visible to the CTA
  float *out = O.xyz + (size_t)cloud * m * 3;
  for (int j = tid; j < m; j += nthreads) {
    const int k = idxs[j];
    const float x = __ldg(dataset + (size_t)k * 3), y = __ldg(dataset + (size_t)k * 3 + 1), z = __ldg(dataset + (size_t)k * 3 + 2);
    if (O.cf) out[j] = x, out[m + j] = y, out[2 * m + j] = z;
    else out[j * 3] = x, out[j * 3 + 1] = y, out[j * 3 + 2] = z;
  }
}

// TB threads, P points per thread (k = tid + i*TB).  WITH_DIST: `data` is the (n,n) distance matrix.
// SMEM_PTS: the cloud also lives in shared memory as float4, so the coordinates of the point just selected — the
// head of every iteration's dependency chain — cost one LDS.128 instead of three global loads.
template <int TB, int P, bool WITH_DIST, bool SMEM_PTS>
__global__ void __launch_bounds__(TB)
fps_kernel(int n, int m, int log2T, const float *__restrict__ data, float *__restrict__ temp,
           FpsOut O) {
  static_assert(P % 2 == 0, "points are processed in packed pairs");
  constexpr int NW = TB / 32;
  extern __shared__ __align__(16) float4 s_pts[];
  __shared__ int s_val[2][32];
  __shared__ uint32_t s_key[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *dataset = data + (size_t)blockIdx.x * (WITH_DIST ? (size_t)n * n : (size_t)n * 3);
  int *idxs = O.idx + (size_t)blockIdx.x * m;

  fps_u64 PX[P / 2], PY[P / 2], PZ[P / 2];  // points 2h and 2h+1 of this thread, packed
  float td[P];
#pragma unroll
  for (int h = 0; h < P / 2; h++) {
    float x[2] = {0.f, 0.f}, y[2] = {0.f, 0.f}, z[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int i = 2 * h + e, k = tid + i * TB;
      if (k < n) {
        if (!WITH_DIST) {
          x[e] = __ldg(dataset + k * 3 + 0);
          y[e] = __ldg(dataset + k * 3 + 1);
          z[e] = __ldg(dataset + k * 3 + 2);
          if (SMEM_PTS) s_pts[k] = make_float4(x[e], y[e], z[e], 0.f);
        }
        td[i] = 1e10f;  // furthest_point_sample.py:30
      } else {
        td[i] = -1.f;  // never selected: compares below every real distance as a signed int
      }
    }
    PX[h] = fps_pack2(x[0], x[1]);
    PY[h] = fps_pack2(y[0], y[1]);
    PZ[h] = fps_pack2(z[0], z[1]);
  }
  // tie-break keys of this thread's points, once (they sit on every iteration's critical path otherwise); for the
  // register-starved 32-points-per-thread shape they are recomputed on demand
  constexpr bool kKeys = P <= 16;
  uint32_t pkey[kKeys ? P : 1];
  if (kKeys) {
#pragma unroll
    for (int i = 0; i < P; i++) pkey[i] = (tid + i * TB < n) ? fps_key(tid + i * TB, log2T) : 0xffffffffu;
  }
  if (warp == 0) {
    s_val[0][lane] = s_val[1][lane] = (int)0x80000000;
    s_key[0][lane] = s_key[1][lane] = 0xffffffffu;
  }
  int old = 0;
  if (tid == 0) idxs[0] = 0;
  __syncthreads();
#ifdef MVP_FPS_TIMING
  long long fps_acc[6] = {0, 0, 0, 0, 0, 0}, fps_last = clock64();
#endif

  for (int j = 1; j < m; j++) {
    float vmax = -1.f;
    if (WITH_DIST) {
      const float *row = dataset + (size_t)old * n;
#pragma unroll
      for (int i = 0; i < P; i++) {
        const int k = tid + i * TB;
        if (k < n) td[i] = fminf(__ldg(row + k), td[i]);
        vmax = fmaxf(vmax, td[i]);
      }
    } else {
      float x1, y1, z1;
      if (SMEM_PTS) {
        const float4 o = s_pts[old];
        x1 = o.x, y1 = o.y, z1 = o.z;
      } else {
        x1 = __ldg(dataset + old * 3 + 0);
        y1 = __ldg(dataset + old * 3 + 1);
        z1 = __ldg(dataset + old * 3 + 2);
      }
      const fps_u64 qx = fps_pack2(x1, x1), qy = fps_pack2(y1, y1), qz = fps_pack2(z1, z1);
#pragma unroll
      for (int h = 0; h < P / 2; h++) {
        float d0, d1;
        fps_unpack2(fps_dist2(PX[h], PY[h], PZ[h], qx, qy, qz), d0, d1);  // point - old (:65-66)
        td[2 * h] = fminf(d0, td[2 * h]);
        td[2 * h + 1] = fminf(d1, td[2 * h + 1]);
        vmax = fps_max3(vmax, td[2 * h], td[2 * h + 1]);
      }
    }
    const int vbits = __float_as_int(vmax);
    MVP_FPS_STAMP(0);
    const int wbits = redux_max_s32(vbits);
    uint32_t key = 0xffffffffu;
    MVP_FPS_STAMP(1);
    if (vbits == wbits) {
#pragma unroll
      for (int i = 0; i < P; i++) {
        const int k = tid + i * TB;
        if (kKeys) {
          if (__float_as_int(td[i]) == wbits) key = min(key, pkey[i]);  // padding: td = -1 never equals a maximum
        } else if (k < n && __float_as_int(td[i]) == wbits) {
          key = min(key, fps_key(k, log2T));
        }
      }
    }
    MVP_FPS_STAMP(2);
    const uint32_t wkey = redux_min_u32(key);
    const int buf = j & 1;
    if (lane == 0) {
      s_val[buf][warp] = wbits;
      s_key[buf][warp] = wkey;
    }
    MVP_FPS_STAMP(3);
    __syncthreads();
    const int v = s_val[buf][lane];  // slots >= NW hold the sentinel
    const uint32_t kk = s_key[buf][lane];
    const int bv = redux_max_s32(v);
    const uint32_t bk = redux_min_u32(v == bv ? kk : 0xffffffffu);
    old = fps_unkey(bk, log2T);
    MVP_FPS_STAMP(4);
    if (tid == 0) idxs[j] = old;
    (void)NW;
  }
#ifdef MVP_FPS_TIMING
  if (tid == 0 && blockIdx.x == 0)
    printf("fps TB %d P %d n %d m %d: cycles/pick  update %lld redux.max %lld keys %lld redux.min+STS %lld bar+LDS+2redux+unkey %lld\n",
           TB, P, n, m, fps_acc[0] / (m - 1), fps_acc[1] / (m - 1), fps_acc[2] / (m - 1), fps_acc[3] / (m - 1),
           fps_acc[4] / (m - 1));
#endif
  if (!WITH_DIST) fps_write_points(O, blockIdx.x, m, idxs, dataset, tid, TB);
  if (temp != nullptr) {
    temp += (size_t)blockIdx.x * n;
#pragma unroll
    for (int i = 0; i < P; i++) {
      const int k = tid + i * TB;
      if (k < n) temp[k] = td[i];
    }
  }
}

// ---- spatially sorted variant (xyz input, clouds of up to 8192 points) ----------------------------------------------------
// Same picks, bit for bit; what changes is how much of the cloud a pick touches.  The CTA first sorts its cloud by the
// cells of a uniform grid (about P points per cell, shared-memory counting sort), so the P points a thread owns are
// neighbours in space and fit a small box.  td[k] = min(td[k], d(k, new)) cannot change when the new point is further
// from the thread's box than the largest td inside it, so such a thread keeps its cached (max, key) and a warp without
// any affected lane skips the update, both `redux` and the key selection.  After the first few dozen picks only the
// neighbourhood of the new point is affected — one or two warps of eight.  The skip test is conservative: box distance
// (monotone in every rounding step of the distance the update would compute) scaled by (1 - 1e-5); NaN compares false,
// i.e. "update".  Tie order: keys are those of the ORIGINAL indices, so ownership of points by threads is irrelevant.
template <int TB, int P>
__global__ void __launch_bounds__(TB)
fps_sorted_kernel(int n, int m, int log2T, const float *__restrict__ data, float *__restrict__ temp,
                  FpsOut O) {
  static_assert(P % 2 == 0, "points are processed in packed pairs");
  extern __shared__ __align__(16) float4 s_pts[];   // [n] by original index | order[n] | hist[cap]
  __shared__ float s_red[6][32];
  __shared__ int s_fin[32], s_warp[32];
  __shared__ GridHdr s_hdr;
  __shared__ int s_val[2][32];
  __shared__ uint32_t s_key[2][32];
  int *s_order = reinterpret_cast<int *>(s_pts + n);
  int *s_hist = s_order + n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *dataset = data + (size_t)blockIdx.x * n * 3;
  int *idxs = O.idx + (size_t)blockIdx.x * m;
  const float inf = __int_as_float(0x7f800000);

  // ---- the cloud into shared memory; bounding box
  float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  int fin = 1;
  for (int k = tid; k < n; k += TB) {
    const float x = __ldg(dataset + k * 3 + 0), y = __ldg(dataset + k * 3 + 1), z = __ldg(dataset + k * 3 + 2);
    s_pts[k] = make_float4(x, y, z, 0.f);
    lo[0] = fminf(lo[0], x), lo[1] = fminf(lo[1], y), lo[2] = fminf(lo[2], z);
    hi[0] = fmaxf(hi[0], x), hi[1] = fmaxf(hi[1], y), hi[2] = fmaxf(hi[2], z);
    fin &= (fabsf(x) <= 3.0e38f && fabsf(y) <= 3.0e38f && fabsf(z) <= 3.0e38f) ? 1 : 0;
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
    }
    fin &= __shfl_xor_sync(0xffffffffu, fin, off);
  }
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) s_red[a][warp] = lo[a], s_red[3 + a][warp] = hi[a];
    s_fin[warp] = fin;
  }
  if (warp == 0) {
    s_val[0][lane] = s_val[1][lane] = (int)0x80000000;
    s_key[0][lane] = s_key[1][lane] = 0xffffffffu;
  }
  __syncthreads();
  const int cap = max(8, n / P);  // about P points per cell: a thread's chunk of the sorted order is one cell's worth
  if (tid == 0) {
    for (int w = 1; w < TB / 32; w++) {
#pragma unroll
      for (int a = 0; a < 3; a++) lo[a] = fminf(lo[a], s_red[a][w]), hi[a] = fmaxf(hi[a], s_red[3 + a][w]);
      fin &= s_fin[w];
    }
    s_hdr = grid_header(lo, hi, fin, cap);
  }
  __syncthreads();
  const GridHdr h = s_hdr;
  const int ncell = h.ncell;
  for (int c = tid; c < ncell; c += TB) s_hist[c] = 0;
  __syncthreads();
  auto cell_of = [&](const float4 &p) {
    if (!h.valid) return 0;
    const int cx = cell_coord((p.x - h.lo[0]) * h.inv_s, h.g[0]);
    const int cy = cell_coord((p.y - h.lo[1]) * h.inv_s, h.g[1]);
    const int cz = cell_coord((p.z - h.lo[2]) * h.inv_s, h.g[2]);
    return (cz * h.g[1] + cy) * h.g[0] + cx;
  };
  for (int k = tid; k < n; k += TB) atomicAdd(&s_hist[cell_of(s_pts[k])], 1);
  __syncthreads();
  {  // exclusive scan of the cell counts -> fill cursors
    const int per = (ncell + TB - 1) / TB;
    const int c0 = min(tid * per, ncell), c1 = min(c0 + per, ncell);
    int sum = 0;
    for (int c = c0; c < c1; c++) sum += s_hist[c];
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = lane < TB / 32 ? s_warp[lane] : 0;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= off) v += u;
      }
      s_warp[lane] = v;
    }
    __syncthreads();
    int run = incl - sum + (warp ? s_warp[warp - 1] : 0);
    for (int c = c0; c < c1; c++) {
      const int cnt = s_hist[c];
      s_hist[c] = run;
      run += cnt;
    }
  }
  __syncthreads();
  for (int k = tid; k < n; k += TB) s_order[atomicAdd(&s_hist[cell_of(s_pts[k])], 1)] = k;
  __syncthreads();

  // ---- this thread's P points: sorted positions [tid * P, tid * P + P)
  fps_u64 PX[P / 2], PY[P / 2], PZ[P / 2];
  float td[P];
  uint32_t pkey[P];
  float blo[3] = {inf, inf, inf}, bhi[3] = {-inf, -inf, -inf};
  uint32_t tkey = 0xffffffffu;
#pragma unroll
  for (int hh = 0; hh < P / 2; hh++) {
    float x[2] = {0.f, 0.f}, y[2] = {0.f, 0.f}, z[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int i = 2 * hh + e, pos = tid * P + i;
      if (pos < n) {
        const int k = s_order[pos];
        const float4 p = s_pts[k];
        x[e] = p.x, y[e] = p.y, z[e] = p.z;
        blo[0] = fminf(blo[0], p.x), blo[1] = fminf(blo[1], p.y), blo[2] = fminf(blo[2], p.z);
        bhi[0] = fmaxf(bhi[0], p.x), bhi[1] = fmaxf(bhi[1], p.y), bhi[2] = fmaxf(bhi[2], p.z);
        td[i] = 1e10f;  // furthest_point_sample.py:30
        pkey[i] = fps_key(k, log2T);
        tkey = min(tkey, pkey[i]);
      } else {
        td[i] = -1.f;  // never selected
        pkey[i] = 0xffffffffu;
      }
    }
    PX[hh] = fps_pack2(x[0], x[1]);
    PY[hh] = fps_pack2(y[0], y[1]);
    PZ[hh] = fps_pack2(z[0], z[1]);
  }
  float tmax = tid * P < n ? 1e10f : -1.f;  // largest running minimum of this thread's points, tkey: its key
  int wbits = 0;
  uint32_t wkey = 0xffffffffu;
  int old = 0;
  if (tid == 0) idxs[0] = 0;

  for (int j = 1; j < m; j++) {
    const float4 o = s_pts[old];
    const float ex = fmaxf(fmaxf(blo[0] - o.x, o.x - bhi[0]), 0.f);
    const float ey = fmaxf(fmaxf(blo[1] - o.y, o.y - bhi[1]), 0.f);
    const float ez = fmaxf(fmaxf(blo[2] - o.z, o.z - bhi[2]), 0.f);
    const float lb2 = sqdist(ex, ey, ez) * (1.f - 1e-5f);
    const bool affected = !(lb2 > tmax);  // NaN anywhere -> update
    if (__any_sync(0xffffffffu, affected) || j == 1) {
      if (affected) {
        const fps_u64 qx = fps_pack2(o.x, o.x), qy = fps_pack2(o.y, o.y), qz = fps_pack2(o.z, o.z);
        float vmax = -1.f;
#pragma unroll
        for (int hh = 0; hh < P / 2; hh++) {
          float d0, d1;
          fps_unpack2(fps_dist2(PX[hh], PY[hh], PZ[hh], qx, qy, qz), d0, d1);  // point - old (:65-66)
          td[2 * hh] = fminf(d0, td[2 * hh]);
          td[2 * hh + 1] = fminf(d1, td[2 * hh + 1]);
          vmax = fps_max3(vmax, td[2 * hh], td[2 * hh + 1]);
        }
        uint32_t key = 0xffffffffu;
        const int vb = __float_as_int(vmax);
#pragma unroll
        for (int i = 0; i < P; i++)
          if (__float_as_int(td[i]) == vb) key = min(key, pkey[i]);  // padding (td = -1, key all ones) changes nothing
        tmax = vmax;
        tkey = key;
      }
      const int vbits = __float_as_int(tmax);
      wbits = redux_max_s32(vbits);
      wkey = redux_min_u32(vbits == wbits ? tkey : 0xffffffffu);
    }
    const int buf = j & 1;
    if (lane == 0) {
      s_val[buf][warp] = wbits;
      s_key[buf][warp] = wkey;
    }
    __syncthreads();
    const int v = s_val[buf][lane];  // slots >= TB / 32 hold the sentinel
    const uint32_t kk = s_key[buf][lane];
    const int bv = redux_max_s32(v);
    const uint32_t bk = redux_min_u32(v == bv ? kk : 0xffffffffu);
    old = fps_unkey(bk, log2T);
    if (tid == 0) idxs[j] = old;
  }
  fps_write_points(O, blockIdx.x, m, idxs, dataset, tid, TB);
  if (temp != nullptr) {
    temp += (size_t)blockIdx.x * n;
#pragma unroll
    for (int i = 0; i < P; i++) {
      const int pos = tid * P + i;
      if (pos < n) temp[s_order[pos]] = td[i];
    }
  }
}

// Fallback for clouds too large for registers (n > 32768): running minima in global `temp`.
template <bool WITH_DIST>
__global__ void __launch_bounds__(1024)
fps_big_kernel(int n, int m, int log2T, const float *__restrict__ data, float *__restrict__ temp,
               FpsOut O) {
  __shared__ int s_val[2][32];
  __shared__ uint32_t s_key[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *dataset = data + (size_t)blockIdx.x * (WITH_DIST ? (size_t)n * n : (size_t)n * 3);
  int *idxs = O.idx + (size_t)blockIdx.x * m;
  temp += (size_t)blockIdx.x * n;
  for (int k = tid; k < n; k += 1024) temp[k] = 1e10f;
  int old = 0;
  if (tid == 0) idxs[0] = 0;
  __syncthreads();
  for (int j = 1; j < m; j++) {
    float x1 = 0, y1 = 0, z1 = 0;
    if (!WITH_DIST) {
      x1 = __ldg(dataset + old * 3 + 0);
      y1 = __ldg(dataset + old * 3 + 1);
      z1 = __ldg(dataset + old * 3 + 2);
    }
    int vbits = __float_as_int(-1.f);
    uint32_t key = 0xffffffffu;
    for (int k = tid; k < n; k += 1024) {
      float d;
      if (WITH_DIST)
        d = __ldg(dataset + (size_t)old * n + k);
      else
        d = sqdist(__ldg(dataset + k * 3 + 0) - x1, __ldg(dataset + k * 3 + 1) - y1,
                   __ldg(dataset + k * 3 + 2) - z1);
      const float d2 = fminf(d, temp[k]);
      temp[k] = d2;
      const int bits = __float_as_int(d2);
      const uint32_t kk = fps_key(k, log2T);
      if (bits > vbits || (bits == vbits && kk < key)) {
        vbits = bits;
        key = kk;
      }
    }
    const int wbits = redux_max_s32(vbits);
    const uint32_t wkey = redux_min_u32(vbits == wbits ? key : 0xffffffffu);
    const int buf = j & 1;
    if (lane == 0) {
      s_val[buf][warp] = wbits;
      s_key[buf][warp] = wkey;
    }
    __syncthreads();
    const int v = s_val[buf][lane];
    const uint32_t kk = s_key[buf][lane];
    const int bv = redux_max_s32(v);
    const uint32_t bk = redux_min_u32(v == bv ? kk : 0xffffffffu);
    old = fps_unkey(bk, log2T);
    if (tid == 0) idxs[j] = old;
  }
  if (!WITH_DIST) fps_write_points(O, blockIdx.x, m, idxs, dataset, tid, 1024);
}

// ---- one cloud split over a CLUSTER of C CTAs (xyz input) ------------------------------------------------------------------
// One CTA per cloud leaves 148 - B SMs idle and makes every pick wait for P distance updates per thread.  Here C CTAs
// share a cloud: CTA r owns the points k = (i C + r) TB + tid, so a pick costs P / C updates per thread; every warp
// then sends its (max, key) slot to ALL C CTAs through distributed shared memory — st.async with mbarrier complete_tx,
// so the data and its arrival signal travel together — and every CTA reduces the C x TB/32 slots redundantly after
// waiting on its OWN mbarrier: no cluster-wide barrier (380 cycles + an L1 flush), one DSMEM hop (~215 cycles) per pick.
// Slots and mbarriers are double-buffered by pick parity: a CTA can be at most one pick ahead of its peers.
// Same picks, bit for bit (keys are those of the original indices).  Used where it measured faster: fps_cluster_size().
template <int TB, int P, int C>
__global__ void __launch_bounds__(TB)
fps_cluster_kernel(int n, int m, int log2T, const float *__restrict__ data, float *__restrict__ temp,
                   FpsOut O) {
  static_assert(P % 2 == 0 && C * (TB / 32) <= 32, "packed pairs; the slots of a pick fit one warp");
  constexpr int NW = TB / 32;
  extern __shared__ __align__(16) float4 s_pts[];
  __shared__ __align__(8) unsigned long long s_slot[2][32];  // (max bits << 32) | key, one per warp of the cluster
  __shared__ __align__(8) uint64_t s_bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int cloud = blockIdx.x / C;
  const float *dataset = data + (size_t)cloud * n * 3;
  int *idxs = O.idx + (size_t)cloud * m;

  for (int k = tid; k < n; k += TB)
    s_pts[k] = make_float4(__ldg(dataset + k * 3 + 0), __ldg(dataset + k * 3 + 1), __ldg(dataset + k * 3 + 2), 0.f);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) s_slot[0][lane] = s_slot[1][lane] = 0x80000000ffffffffULL;  // sentinel: below every real maximum
  __syncthreads();
  fps_u64 PX[P / 2], PY[P / 2], PZ[P / 2];
  float td[P];
  uint32_t pkey[P];
#pragma unroll
  for (int h = 0; h < P / 2; h++) {
    float x[2] = {0.f, 0.f}, y[2] = {0.f, 0.f}, z[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int i = 2 * h + e, k = (i * C + (int)rank) * TB + tid;
      if (k < n) {
        const float4 q = s_pts[k];
        x[e] = q.x, y[e] = q.y, z[e] = q.z;
        td[i] = 1e10f;
        pkey[i] = fps_key(k, log2T);
      } else {
        td[i] = -1.f;
        pkey[i] = 0xffffffffu;
      }
    }
    PX[h] = fps_pack2(x[0], x[1]), PY[h] = fps_pack2(y[0], y[1]), PZ[h] = fps_pack2(z[0], z[1]);
  }
  // remote addresses of this warp's slot and of the mbarrier in every CTA of the cluster
  uint32_t r_slot[2][C], r_bar[2][C];
#pragma unroll
  for (int bf = 0; bf < 2; bf++) {
#pragma unroll
    for (int c = 0; c < C; c++) {
      const uint32_t ls = (uint32_t)__cvta_generic_to_shared(&s_slot[bf][rank * NW + warp]);
      const uint32_t lb = (uint32_t)__cvta_generic_to_shared(&s_bar[bf]);
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r_slot[bf][c]) : "r"(ls), "r"(c));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r_bar[bf][c]) : "r"(lb), "r"(c));
    }
  }
  // every CTA's barriers and sentinels are initialised before anyone's first remote store lands
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  int old = 0;
  if (tid == 0 && rank == 0) idxs[0] = 0;

  for (int j = 1; j < m; j++) {
    const int buf = j & 1;
    if (tid == 0)  // this pick's slots: C * NW stores of 8 bytes are expected on the local barrier
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bar[buf])),
                   "r"(C * NW * 8)
                   : "memory");
    const float4 o = s_pts[old];
    const fps_u64 qx = fps_pack2(o.x, o.x), qy = fps_pack2(o.y, o.y), qz = fps_pack2(o.z, o.z);
    float vmax = -1.f;
#pragma unroll
    for (int h = 0; h < P / 2; h++) {
      float d0, d1;
      fps_unpack2(fps_dist2(PX[h], PY[h], PZ[h], qx, qy, qz), d0, d1);
      td[2 * h] = fminf(d0, td[2 * h]);
      td[2 * h + 1] = fminf(d1, td[2 * h + 1]);
      vmax = fps_max3(vmax, td[2 * h], td[2 * h + 1]);
    }
    const int vbits = __float_as_int(vmax);
    const int wbits = redux_max_s32(vbits);
    uint32_t key = 0xffffffffu;
    if (vbits == wbits) {
#pragma unroll
      for (int i = 0; i < P; i++)
        if (__float_as_int(td[i]) == wbits) key = min(key, pkey[i]);
    }
    const uint32_t wkey = redux_min_u32(key);
    if (lane < C) {  // lane c sends this warp's slot to CTA c
      const unsigned long long v = ((unsigned long long)(uint32_t)wbits << 32) | wkey;
      uint32_t dst = r_slot[buf][0], bar = r_bar[buf][0];
#pragma unroll
      for (int c = 1; c < C; c++)
        if (lane == c) dst = r_slot[buf][c], bar = r_bar[buf][c];
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst), "l"(v), "r"(bar)
                   : "memory");
    }
    {  // wait for all C * NW slots of this pick
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[buf]);
      const uint32_t parity = (uint32_t)((j - 1) >> 1) & 1u;  // picks 1, 2 are the first use of either barrier
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "FPSW_%=:\n\t"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
          "@p bra FPSD_%=;\n\t"
          "bra FPSW_%=;\n\t"
          "FPSD_%=:\n\t}" ::"r"(bar),
          "r"(parity)
          : "memory");
    }
    const unsigned long long sv = s_slot[buf][lane];  // slots >= C * NW hold the sentinel
    const int v = (int)(uint32_t)(sv >> 32);
    const uint32_t kk = (uint32_t)sv;
    const int bv = redux_max_s32(v);
    const uint32_t bk = redux_min_u32(v == bv ? kk : 0xffffffffu);
    old = fps_unkey(bk, log2T);
    if (tid == 0 && rank == 0) idxs[j] = old;
  }
  if (rank == 0) fps_write_points(O, cloud, m, idxs, dataset, tid, TB);
  if (temp != nullptr) {
    temp += (size_t)cloud * n;
#pragma unroll
    for (int i = 0; i < P; i++) {
      const int k = (i * C + (int)rank) * TB + tid;
      if (k < n) temp[k] = td[i];
    }
  }
  // no CTA leaves while a peer may still address its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int TB, int P, int C>
static int fps_cluster_launch(int b, int n, int m, int log2T, const float *data, float *temp, FpsOut idx, cudaStream_t s) {
  const size_t smem = (size_t)n * sizeof(float4);
  static size_t granted[kMaxDevices];
  int rc = grant_dyn_smem(fps_cluster_kernel<TB, P, C>, smem > 40 * 1024 ? (size_t)TB * P * C * sizeof(float4) : smem, granted);
  if (rc) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(b * C);
  cfg.blockDim = dim3(TB);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, fps_cluster_kernel<TB, P, C>, n, m, log2T, data, temp, idx);
  return e == cudaSuccess ? MVP_OK : (int)e;
}

// Cluster size for a problem: MVP_FPS_CLUSTER=0|2|4 forces it (A/B timing).  Measured (tools/fps_sizes.py, ns per pick,
// 1 / 2 / 4 CTAs per cloud): 32x2048 242 / 302 / 295, 64x3072 304 / 320 / 310, 64x1536 249 / 320 / 308, 64x768 231 /
// 286 / 288, 16x8192 458 (sorted kernel) / 464 / 418 — the DSMEM hop costs more than the shorter update saves until a
// thread would own sixteen points, so only clouds above 4096 points are split, four ways, while the clusters fit the
// machine in one wave.
static int fps_cluster_size(int b, int n) {
  static const int forced = [] {
    const char *e = getenv("MVP_FPS_CLUSTER");
    const int v = e ? atoi(e) : -1;
    return (v == 0 || v == 2 || v == 4) ? v : -1;
  }();
  if (n < 512 || n > 8192) return 0;
  if (forced >= 0) return forced;
  return (n > 4096 && b * 4 <= kNumSMs) ? 4 : 0;
}

// furthest_point_sample_cuda.cu:11-15 — evaluated with the same double expression so that the tie order
// matches the block size the reference would have launched.
static int ref_block_size(int work_size) {
  const int pow_2 = (int)(std::log(static_cast<double>(work_size)) / std::log(2.0));
  int t = 1 << pow_2;
  if (t > 1024) t = 1024;
  if (t < 1) t = 1;
  return t;
}

static bool fps_use_sorted() {  // MVP_FPS_SORTED=0: the unsorted kernel (A/B timing)
  static const bool on = [] {
    const char *e = getenv("MVP_FPS_SORTED");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <bool WD, int TB, int P>
static void fps_launch_one(int b, int n, int m, int log2T, const float *data, float *temp, FpsOut idx,
                           cudaStream_t s) {
  if constexpr (!WD && TB * P <= 8192) {
    // measured (tools/fps_sizes.py): 457 against 601 ns per pick at n = 8192, but 318 against 242 at n = 2048 — with
    // 8 points per thread the boxes are a sizeable part of a small cloud, nearly every warp keeps an affected lane,
    // and the test only lengthens the dependent chain of a pick
    if (fps_use_sorted() && n > 4096) {
      const size_t smem = (size_t)n * (sizeof(float4) + sizeof(int)) + sizeof(int) * (size_t)std::max(8, n / P);
      static size_t granted[kMaxDevices];
      const size_t want = (size_t)TB * P * (sizeof(float4) + sizeof(int)) + sizeof(int) * (size_t)std::max(8, TB);
      if (grant_dyn_smem(fps_sorted_kernel<TB, P>, smem > 40 * 1024 ? want : smem, granted)) return;  // launch_status() reports it
      fps_sorted_kernel<TB, P><<<b, TB, smem, s>>>(n, m, log2T, data, temp, idx);
      return;
    }
  }
  // the shared-memory copy of the cloud: 16 B per point, for clouds up to 8192 points (128 KB)
  constexpr bool kSmem = !WD && TB * P <= 8192;
  const size_t smem = kSmem ? (size_t)n * sizeof(float4) : 0;
  {
    static size_t granted[kMaxDevices];
    if (grant_dyn_smem(fps_kernel<TB, P, WD, kSmem>, smem > 40 * 1024 ? TB * P * sizeof(float4) : smem, granted))
      return;  // launch_status() reports it
  }
  fps_kernel<TB, P, WD, kSmem><<<b, TB, smem, s>>>(n, m, log2T, data, temp, idx);
}

template <bool WD>
static int fps_dispatch(int b, int n, int m, const float *data, float *temp, FpsOut idx, cudaStream_t s) {
  if (b < 0 || n < 0 || m < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || m == 0) return MVP_OK;  // reference kernel returns immediately when m <= 0 (:34)
  if (n == 0 || !data || !idx.idx) return MVP_ERR_INVALID_ARGUMENT;
  const int T = ref_block_size(n);
  int log2T = 0;
  while ((1 << log2T) < T) log2T++;
  if (!WD && fps_cluster_size(b, n)) {
    const int c = fps_cluster_size(b, n);
    int rc;
    // TB * P * C >= n
    if (c == 2) {
      if (n <= 1024) rc = fps_cluster_launch<128, 4, 2>(b, n, m, log2T, data, temp, idx, s);
      else if (n <= 2048) rc = fps_cluster_launch<256, 4, 2>(b, n, m, log2T, data, temp, idx, s);
      else if (n <= 3072) rc = fps_cluster_launch<256, 6, 2>(b, n, m, log2T, data, temp, idx, s);
      else if (n <= 4096) rc = fps_cluster_launch<256, 8, 2>(b, n, m, log2T, data, temp, idx, s);
      else rc = fps_cluster_launch<512, 8, 2>(b, n, m, log2T, data, temp, idx, s);
    } else {
      if (n <= 1024) rc = fps_cluster_launch<128, 2, 4>(b, n, m, log2T, data, temp, idx, s);
      else if (n <= 2048) rc = fps_cluster_launch<128, 4, 4>(b, n, m, log2T, data, temp, idx, s);
      else if (n <= 3072) rc = fps_cluster_launch<128, 6, 4>(b, n, m, log2T, data, temp, idx, s);
      else if (n <= 4096) rc = fps_cluster_launch<256, 4, 4>(b, n, m, log2T, data, temp, idx, s);
      else rc = fps_cluster_launch<256, 8, 4>(b, n, m, log2T, data, temp, idx, s);
    }
    count_launch();
    return rc ? rc : launch_status();
  }
  // (TB, P) with TB*P >= n: at most MVP_FPS_PMAX points per thread while warps are left, then grow P.
  if (n > 4096 && n <= 512 * 16) fps_launch_one<WD, 512, 16>(b, n, m, log2T, data, temp, idx, s);  // measured best
  else
#if MVP_FPS_PMAX >= 16
  if (n <= 128 * 4) fps_launch_one<WD, 128, 4>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 128 * 8) fps_launch_one<WD, 128, 8>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 128 * 16) fps_launch_one<WD, 128, 16>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 256 * 12) fps_launch_one<WD, 256, 12>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 256 * 16) fps_launch_one<WD, 256, 16>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 512 * 16) fps_launch_one<WD, 512, 16>(b, n, m, log2T, data, temp, idx, s);
#elif MVP_FPS_PMAX >= 8
  if (n <= 128 * 4) fps_launch_one<WD, 128, 4>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 128 * 8) fps_launch_one<WD, 128, 8>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 256 * 8) fps_launch_one<WD, 256, 8>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 512 * 6) fps_launch_one<WD, 512, 6>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 512 * 8) fps_launch_one<WD, 512, 8>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 1024 * 8) fps_launch_one<WD, 1024, 8>(b, n, m, log2T, data, temp, idx, s);
#else
  if (n <= 128 * 4) fps_launch_one<WD, 128, 4>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 256 * 4) fps_launch_one<WD, 256, 4>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 512 * 4) fps_launch_one<WD, 512, 4>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 1024 * 4) fps_launch_one<WD, 1024, 4>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 1024 * 8) fps_launch_one<WD, 1024, 8>(b, n, m, log2T, data, temp, idx, s);
#endif
  else if (n <= 1024 * 16) fps_launch_one<WD, 1024, 16>(b, n, m, log2T, data, temp, idx, s);
  else if (n <= 1024 * 32) fps_launch_one<WD, 1024, 32>(b, n, m, log2T, data, temp, idx, s);
  else {
    if (!temp) return MVP_ERR_WORKSPACE;  // the big-cloud path needs the reference's `temp` scratch
    fps_big_kernel<WD><<<b, 1024, 0, s>>>(n, m, log2T, data, temp, idx);
  }
  count_launch();
  return launch_status();
}

}  // namespace mvp

MVP_API int mvp_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                        mvp_stream_t stream) {
  return mvp::fps_dispatch<false>(b, n, m, xyz, temp, mvp::FpsOut{idx, nullptr, 0}, (cudaStream_t)stream);
}

MVP_API int mvp_furthest_point_sampling_with_dist(int b, int n, int m, const float *dist, float *temp,
                                                  int *idx, mvp_stream_t stream) {
  return mvp::fps_dispatch<true>(b, n, m, dist, temp, mvp::FpsOut{idx, nullptr, 0}, (cudaStream_t)stream);
}

// furthest_point_sample + gather_points of the sampled coordinates in one launch (SURVEY.md §8(f) row 2:
// completion/model_utils.py:91-93, vrcnet.py:451): idx (b, m) as above and sampled_xyz (b, m, 3), or (b, 3, m) — the
// layout gather_points returns — when channels_first != 0.
MVP_API int mvp_furthest_point_sampling_gather(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                               float *sampled_xyz, int channels_first, mvp_stream_t stream) {
  if (!sampled_xyz && b > 0 && m > 0) return MVP_ERR_INVALID_ARGUMENT;
  return mvp::fps_dispatch<false>(b, n, m, xyz, temp, mvp::FpsOut{idx, sampled_xyz, channels_first ? 1 : 0}, (cudaStream_t)stream);
}
