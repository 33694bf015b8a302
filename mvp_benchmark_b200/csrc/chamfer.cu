// Chamfer distance for sm_100a — generic one-direction kernel + fused backward.
//
// Replaces NmDistanceKernel / NmDistanceGradKernel of the reference
// (utils/metrics/CD/chamfer3D/chamfer3D.cu:12-134,155-174).  Semantics kept bit-for-bit:
//   d = fma(dz,dz, fma(dx,dx, dy*dy)) with dx = target - query (chamfer3D.cu:32-35 as contracted by nvcc),
//   argmin = lowest target index among equal minima (strict `<` scanning upward, :36,:126).
//
// This file holds the GENERIC path (any n, m >= 1): one launch per direction, R queries per thread in
// registers, targets staged through shared memory.  The large-cloud fast path (both directions from one
// evaluation of each pair, packed fp32x2 math, TMA-staged tiles) lives in chamfer_fused.cu and is chosen
// by mvp_chamfer_forward when the shapes allow.
#include <cstdlib>

#include "common.cuh"

namespace mvp {

constexpr int kChThreads = 256;
constexpr int kChTile = 1024;  // targets per shared-memory tile (AoS, 12 KB)

template <int R>
__device__ __forceinline__ void chamfer_dir_body(int chunk, int b, int n, int m, const float *__restrict__ xyz,
                                                 const float *__restrict__ xyz2, float *__restrict__ dist,
                                                 int *__restrict__ idx) {
  __shared__ __align__(16) float tile[kChTile * 3];
  const int tid = threadIdx.x;
  const int qbase = chunk * (kChThreads * R);
  const int nq = n;
  if (qbase >= nq) return;
  const float *q = xyz + (size_t)b * n * 3;
  const float *t = xyz2 + (size_t)b * m * 3;

  float qx[R], qy[R], qz[R], best[R];
  int bi[R], qi[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int j = qbase + r * kChThreads + tid;
    const int jj = j < nq ? j : nq - 1;
    qi[r] = jj;
    qx[r] = __ldg(q + jj * 3 + 0);
    qy[r] = __ldg(q + jj * 3 + 1);
    qz[r] = __ldg(q + jj * 3 + 2);
    best[r] = __int_as_float(0x7f800000);  // +inf
    bi[r] = 0;
  }

  for (int k2 = 0; k2 < m; k2 += kChTile) {
    const int cnt = min(kChTile, m - k2);
    const int cnt4 = (cnt + 3) & ~3;
    for (int i = tid; i < cnt4 * 3; i += kChThreads)
      tile[i] = i < cnt * 3 ? __ldg(t + (size_t)k2 * 3 + i) : __int_as_float(0x7f800000);
    __syncthreads();
    const float4 *t4 = reinterpret_cast<const float4 *>(tile);
#pragma unroll 2
    for (int k = 0; k < cnt4; k += 4) {
      const float4 a = t4[(k >> 2) * 3 + 0];  // x0 y0 z0 x1
      const float4 c = t4[(k >> 2) * 3 + 1];  // y1 z1 x2 y2
      const float4 e = t4[(k >> 2) * 3 + 2];  // z2 x3 y3 z3
      const int kk = k2 + k;
#pragma unroll
      for (int r = 0; r < R; r++) {
        const float d0 = sqdist(a.x - qx[r], a.y - qy[r], a.z - qz[r]);
        const float d1 = sqdist(a.w - qx[r], c.x - qy[r], c.y - qz[r]);
        const float d2 = sqdist(c.z - qx[r], c.w - qy[r], e.x - qz[r]);
        const float d3 = sqdist(e.y - qx[r], e.z - qy[r], e.w - qz[r]);
        if (d0 < best[r]) { best[r] = d0; bi[r] = kk; }
        if (d1 < best[r]) { best[r] = d1; bi[r] = kk + 1; }
        if (d2 < best[r]) { best[r] = d2; bi[r] = kk + 2; }
        if (d3 < best[r]) { best[r] = d3; bi[r] = kk + 3; }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int j = qbase + r * kChThreads + tid;
    if (j < nq) {
      dist[(size_t)b * n + qi[r]] = best[r];
      idx[(size_t)b * n + qi[r]] = bi[r];
    }
  }
}

template <int R>
__global__ void __launch_bounds__(kChThreads)
chamfer_dir_kernel(int n, int m, const float *__restrict__ xyz, const float *__restrict__ xyz2,
                   float *__restrict__ dist, int *__restrict__ idx) {
  chamfer_dir_body<R>(blockIdx.x, blockIdx.y, n, m, xyz, xyz2, dist, idx);
}

int chamfer_dir_launch(int b, int n, int m, const float *xyz, const float *xyz2, float *dist, int *idx,
                       cudaStream_t s) {
  // Pick R so that the grid has at least ~2 CTAs per SM where the problem allows it.
  auto ctas = [&](int R) { return (long long)b * ((n + kChThreads * R - 1) / (kChThreads * R)); };
  for (int by = 0; by < b; by += 65535) {
    const int bb = min(65535, b - by);
    const float *q = xyz + (size_t)by * n * 3;
    const float *t = xyz2 + (size_t)by * m * 3;
    float *d = dist + (size_t)by * n;
    int *ix = idx + (size_t)by * n;
    if (ctas(4) >= 2 * kNumSMs) {
      dim3 grid((n + kChThreads * 4 - 1) / (kChThreads * 4), bb);
      chamfer_dir_kernel<4><<<grid, kChThreads, 0, s>>>(n, m, q, t, d, ix);
    } else if (ctas(2) >= 2 * kNumSMs) {
      dim3 grid((n + kChThreads * 2 - 1) / (kChThreads * 2), bb);
      chamfer_dir_kernel<2><<<grid, kChThreads, 0, s>>>(n, m, q, t, d, ix);
    } else {
      dim3 grid((n + kChThreads - 1) / kChThreads, bb);
      chamfer_dir_kernel<1><<<grid, kChThreads, 0, s>>>(n, m, q, t, d, ix);
    }
    count_launch();
  }
  return launch_status();
}

// Backward (chamfer3D.cu:155-174: g = grad*2; six float atomics per point into pre-zeroed gradients).  Point i of
// one cloud contributes  v = g (p_i - q_nn(i))  to its OWN gradient row and -v to its neighbour's row in the OTHER
// cloud.  Every row receives exactly one own contribution, so pass 1 WRITES it with a plain coalesced store (which
// is also what zero-fills the gradients) and pass 2 adds the scattered halves with red.global.add.f32: half the
// atomics of the reference and no memset.  One launch each for both directions: thread i < b*n handles point i of
// xyz1, the rest handle points of xyz2.
template <bool kScatter>
__global__ void __launch_bounds__(256)
chamfer_grad_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                    const float *__restrict__ gd1, const float *__restrict__ gd2,
                    const int *__restrict__ idx1, const int *__restrict__ idx2, float *gx1, float *gx2) {
  const long long total1 = (long long)b * n, total = total1 + (long long)b * m;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const bool first = i < total1;
    const long long p = first ? i : i - total1;
    const int na = first ? n : m, nb = first ? m : n;
    const long long cloud = p / na;
    const float *A = first ? xyz1 : xyz2;
    const float *Bp = first ? xyz2 : xyz1;
    float *GA = first ? gx1 : gx2;
    float *GB = first ? gx2 : gx1;
    const int j2 = __ldg((first ? idx1 : idx2) + p);
    const float g = __ldg((first ? gd1 : gd2) + p) * 2;
    const float x1 = __ldg(A + p * 3 + 0), y1 = __ldg(A + p * 3 + 1), z1 = __ldg(A + p * 3 + 2);
    const long long o = (cloud * nb + j2) * 3;
    const float x2 = __ldg(Bp + o + 0), y2 = __ldg(Bp + o + 1), z2 = __ldg(Bp + o + 2);
    const float vx = g * (x1 - x2), vy = g * (y1 - y2), vz = g * (z1 - z2);
    if (!kScatter) {
      GA[p * 3 + 0] = vx;
      GA[p * 3 + 1] = vy;
      GA[p * 3 + 2] = vz;
    } else {
      atomicAdd(GB + o + 0, -vx);
      atomicAdd(GB + o + 1, -vy);
      atomicAdd(GB + o + 2, -vz);
    }
  }
}

// The same two passes with four consecutive points per thread (n and m multiples of 4, 16-byte aligned tensors):
// coordinates, indices and upstream gradients arrive as 128-bit loads, the own halves leave as three 128-bit
// stores, and the scattered halves as ONE 64-bit + one 32-bit reduction per point instead of three 32-bit ones
// (red.global.add.v2.f32 on whichever pair of the 12-byte row is 8-byte aligned).
// MODE 0: own halves, plain stores.  MODE 1: scattered halves, reductions.  MODE 2: both in ONE pass onto gradients the
// launcher has zeroed — the own halves as coalesced 128-bit reductions — so that coordinates, indices and upstream
// gradients are read once instead of twice.
template <int MODE>
__global__ void __launch_bounds__(256)
chamfer_grad4_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                     const float *__restrict__ gd1, const float *__restrict__ gd2,
                     const int *__restrict__ idx1, const int *__restrict__ idx2, float *gx1, float *gx2) {
  const long long total1 = (long long)b * n / 4, total = total1 + (long long)b * m / 4;  // groups of 4 points
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total;
       q += (long long)gridDim.x * blockDim.x) {
    const bool first = q < total1;
    const long long p = (first ? q : q - total1) * 4;
    const int na = first ? n : m, nb = first ? m : n;
    const long long cloud = p / na;  // na % 4 == 0: the four points share a cloud
    const float *A = first ? xyz1 : xyz2;
    const float *Bp = (first ? xyz2 : xyz1) + cloud * nb * 3;
    float *GA = first ? gx1 : gx2;
    float *GB = (first ? gx2 : gx1) + cloud * nb * 3;
    const int4 j4 = __ldg(reinterpret_cast<const int4 *>((first ? idx1 : idx2) + p));
    const float4 g4 = __ldg(reinterpret_cast<const float4 *>((first ? gd1 : gd2) + p));
    const float4 a0 = __ldg(reinterpret_cast<const float4 *>(A + p * 3));
    const float4 a1 = __ldg(reinterpret_cast<const float4 *>(A + p * 3) + 1);
    const float4 a2 = __ldg(reinterpret_cast<const float4 *>(A + p * 3) + 2);
    const float ax[4] = {a0.x, a0.w, a1.z, a2.y}, ay[4] = {a0.y, a1.x, a1.w, a2.z}, az[4] = {a0.z, a1.y, a2.x, a2.w};
    const int jj[4] = {j4.x, j4.y, j4.z, j4.w};
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float vx[4], vy[4], vz[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const float *t = Bp + (long long)jj[e] * 3;
      const float g = gg[e] * 2;
      vx[e] = g * (ax[e] - __ldg(t + 0));
      vy[e] = g * (ay[e] - __ldg(t + 1));
      vz[e] = g * (az[e] - __ldg(t + 2));
    }
    if (MODE == 0) {
      float4 *o = reinterpret_cast<float4 *>(GA + p * 3);
      o[0] = make_float4(vx[0], vy[0], vz[0], vx[1]);
      o[1] = make_float4(vy[1], vz[1], vx[2], vy[2]);
      o[2] = make_float4(vz[2], vx[3], vy[3], vz[3]);
    }
    if (MODE == 2) {
      float *o = GA + p * 3;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(vx[0]), "f"(vy[0]), "f"(vz[0]), "f"(vx[1]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4), "f"(vy[1]), "f"(vz[1]), "f"(vx[2]), "f"(vy[2]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 8), "f"(vz[2]), "f"(vx[3]), "f"(vy[3]), "f"(vz[3]) : "memory");
    }
    if (MODE != 0) {
#pragma unroll
      for (int e = 0; e < 4; e++) {
        float *t = GB + (long long)jj[e] * 3;
        // A 12-byte row sits at offset 0, 4, 8 or 12 of a 16-byte chunk.  Offsets 0 and 4 lie inside ONE chunk: a
        // single 128-bit reduction covers the row, its fourth lane adding +0 to the neighbouring row's element
        // (exact: x + 0 = x; only -0 becomes +0) — where that neighbour exists inside this cloud's rows.  Offsets 8
        // and 12 straddle two chunks: one 64-bit and one 32-bit reduction.
        const unsigned off = (unsigned)(reinterpret_cast<uintptr_t>(t) & 15);
        if (off == 0 && jj[e] + 1 < nb) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(t), "f"(-vx[e]), "f"(-vy[e]), "f"(-vz[e]),
                       "f"(0.f) : "memory");
        } else if (off == 4 && jj[e] > 0) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(t - 1), "f"(0.f), "f"(-vx[e]), "f"(-vy[e]),
                       "f"(-vz[e]) : "memory");
        } else if ((off & 7) == 0) {
          asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(t), "f"(-vx[e]), "f"(-vy[e]) : "memory");
          atomicAdd(t + 2, -vz[e]);
        } else {
          atomicAdd(t, -vx[e]);
          asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(t + 1), "f"(-vy[e]), "f"(-vz[e]) : "memory");
        }
      }
    }
  }
}


// ---- backward without global atomics: every gradient row is SUMMED by one thread ------------------------------------
// gradxyz_s[j] = 2 g_s[j] (p_j - q_idx_s[j])  -  sum over { i : idx_o[i] == j } of 2 g_o[i] (q_i - p_j)
// (chamfer3D.cu:155-174 accumulates the same terms with six float atomics per point into memset gradients).  A CTA
// owns the rows j in [j0, j0 + R) of one side of one cloud.  It reads the OTHER side's index array once into
// registers (coalesced), transposes the entries that point into its range with a counting sort in shared memory
// (native integer atomics: histogram, block scan, scatter of 16-bit source indices), and then each thread produces
// whole rows: own term + the listed terms, one coalesced 12-byte store per row.  No float atomics, no memset, xyz /
// idx / grad read once from HBM (the gathers hit L2).  Lists of up to kCsrOrdered entries are summed in ascending
// source order (deterministic); longer ones in arrival order.  MEASURED SLOWER than the reduction kernels at the headline
// size (40 us against 32 us at B = 32, N = M = 16384: shared-memory atomics cost 2 cycles per lane, and this kernel needs
// two per point where the reductions need 1.5 vector REDG) — so it is the opt-in MVP_CHAMFER_BWD_SUMMED, for callers
// who want gradients without float atomics, not the default.
constexpr int kCsrThreads = 512;
constexpr int kCsrRegs = 32;            // other-side indices per thread: clouds of up to 16384 points
constexpr int kCsrMaxPts = kCsrThreads * kCsrRegs;
constexpr int kCsrOrdered = 8;

__global__ void __launch_bounds__(kCsrThreads)
chamfer_grad_csr_kernel(int n, int m, int R, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                        const float *__restrict__ gd1, const float *__restrict__ gd2, const int *__restrict__ idx1,
                        const int *__restrict__ idx2, float *__restrict__ gx1, float *__restrict__ gx2) {
  extern __shared__ __align__(16) int csr_smem[];
  __shared__ int s_warp[kCsrThreads / 32];
  const int side = blockIdx.y, cloud = blockIdx.z;
  const int nt = side ? m : n, no = side ? n : m;  // rows produced here / entries of the other side's index
  const int j0 = blockIdx.x * R;
  if (j0 >= nt) return;
  const int rows = min(R, nt - j0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int *hist = csr_smem;                                                       // [R + 1]
  unsigned short *perm = reinterpret_cast<unsigned short *>(csr_smem + R + 1);  // [no]
  const float *Ps = (side ? xyz2 : xyz1) + (size_t)cloud * nt * 3;
  const float *Po = (side ? xyz1 : xyz2) + (size_t)cloud * no * 3;
  const float *Gs = (side ? gd2 : gd1) + (size_t)cloud * nt;
  const float *Go = (side ? gd1 : gd2) + (size_t)cloud * no;
  const int *Is = (side ? idx2 : idx1) + (size_t)cloud * nt;
  const int *Io = (side ? idx1 : idx2) + (size_t)cloud * no;
  float *Out = (side ? gx2 : gx1) + (size_t)cloud * nt * 3;

  // the other side's neighbours, relative to this CTA's range (anything outside it becomes a huge unsigned)
  int rel[kCsrRegs];
#pragma unroll
  for (int k = 0; k < kCsrRegs; k++) {
    const int i = k * kCsrThreads + tid;
    rel[k] = i < no ? __ldg(Io + i) - j0 : -1;
  }
  for (int c = tid; c <= R; c += kCsrThreads) hist[c] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kCsrRegs; k++)
    if ((unsigned)rel[k] < (unsigned)rows) atomicAdd(&hist[rel[k]], 1);
  __syncthreads();
  // exclusive scan of hist[0, R) in place (R is a multiple of kCsrThreads: a contiguous chunk per thread)
  const int per = R / kCsrThreads;
  int sum = 0;
  for (int c = 0; c < per; c++) sum += hist[tid * per + c];
  int incl = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = lane < kCsrThreads / 32 ? s_warp[lane] : 0;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, off);
      if (lane >= off) v += u;
    }
    if (lane < kCsrThreads / 32) s_warp[lane] = v;
  }
  __syncthreads();
  int run = incl - sum + (warp ? s_warp[warp - 1] : 0);
  for (int c = 0; c < per; c++) {
    const int cnt = hist[tid * per + c];
    hist[tid * per + c] = run;  // fill cursor; after the scatter it is the END of row c's list
    run += cnt;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kCsrRegs; k++)
    if ((unsigned)rel[k] < (unsigned)rows) perm[atomicAdd(&hist[rel[k]], 1)] = (unsigned short)(k * kCsrThreads + tid);
  __syncthreads();

  // One row per thread and iteration.  The loads of a row form a chain three deep (own index -> neighbour's
  // coordinates; list bounds -> list entries -> their gradients and coordinates), so the inputs of the NEXT row are
  // requested before the current row's chain is walked, and the first four list entries (lists average one entry)
  // are gathered together and then summed in ascending source order; longer lists continue one entry at a time.
  struct Row {
    float px, py, pz, g;
    int q, e0, e1;
  };
  auto load_row = [&](int r) {
    Row w;
    const int j = j0 + r;
    w.px = __ldg(Ps + (size_t)j * 3 + 0), w.py = __ldg(Ps + (size_t)j * 3 + 1), w.pz = __ldg(Ps + (size_t)j * 3 + 2);
    w.q = __ldg(Is + j);
    w.g = __ldg(Gs + j) * 2;
    w.e0 = r ? hist[r - 1] : 0;
    w.e1 = hist[r];
    return w;
  };
  Row cur;
  if (tid < rows) cur = load_row(tid);
  for (int r = tid; r < rows; r += kCsrThreads) {
    const Row w = cur;
    if (r + kCsrThreads < rows) cur = load_row(r + kCsrThreads);
    const int len = w.e1 - w.e0;
    int ent[4];
#pragma unroll
    for (int t = 0; t < 4; t++) ent[t] = t < len ? (int)perm[w.e0 + t] : 0x7fffffff;
    if (len > 4 && len <= kCsrOrdered) {  // keep the four smallest here, in order; the loop below continues above them
#pragma unroll 1
      for (int e = w.e0 + 4; e < w.e1; e++) {
        int v = perm[e];
#pragma unroll
        for (int t = 0; t < 4; t++)
          if (v < ent[t]) { const int o = ent[t]; ent[t] = v; v = o; }
      }
    }
    // ascending order of the four (sorting network; absent entries are INT_MAX and sink to the end)
    auto cswap = [](int &a, int &c) { const int lo = min(a, c), hi = max(a, c); a = lo; c = hi; };
    cswap(ent[0], ent[1]); cswap(ent[2], ent[3]); cswap(ent[0], ent[2]); cswap(ent[1], ent[3]); cswap(ent[1], ent[2]);
    float ox = __ldg(Po + (size_t)w.q * 3 + 0), oy = __ldg(Po + (size_t)w.q * 3 + 1), oz = __ldg(Po + (size_t)w.q * 3 + 2);
    float gi[4], qx[4], qy[4], qz[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int i = t < len ? ent[t] : 0;  // a valid address for absent entries; their terms are not added
      gi[t] = __ldg(Go + i) * 2;
      qx[t] = __ldg(Po + (size_t)i * 3 + 0), qy[t] = __ldg(Po + (size_t)i * 3 + 1), qz[t] = __ldg(Po + (size_t)i * 3 + 2);
    }
    float ax = w.g * (w.px - ox), ay = w.g * (w.py - oy), az = w.g * (w.pz - oz);
#pragma unroll
    for (int t = 0; t < 4; t++)
      if (t < len) {
        ax -= gi[t] * (qx[t] - w.px);
        ay -= gi[t] * (qy[t] - w.py);
        az -= gi[t] * (qz[t] - w.pz);
      }
    if (len > 4) {
      auto add = [&](int i) {
        const float g2 = __ldg(Go + i) * 2;
        ax -= g2 * (__ldg(Po + (size_t)i * 3 + 0) - w.px);
        ay -= g2 * (__ldg(Po + (size_t)i * 3 + 1) - w.py);
        az -= g2 * (__ldg(Po + (size_t)i * 3 + 2) - w.pz);
      };
      if (len <= kCsrOrdered) {
        int last = ent[3];
        for (int t = 4; t < len; t++) {  // the smallest entry above the previous one
          int nxt = 0x7fffffff;
          for (int e = w.e0; e < w.e1; e++) {
            const int i = perm[e];
            if (i > last && i < nxt) nxt = i;
          }
          add(nxt);
          last = nxt;
        }
      } else {  // long list (many queries share one neighbour): arrival order, minus the four already added
        for (int e = w.e0; e < w.e1; e++) {
          const int i = perm[e];
          if (i != ent[0] && i != ent[1] && i != ent[2] && i != ent[3]) add(i);
        }
      }
    }
    const int j = j0 + r;
    Out[(size_t)j * 3 + 0] = ax;
    Out[(size_t)j * 3 + 1] = ay;
    Out[(size_t)j * 3 + 2] = az;
  }
}

// rows per CTA: a multiple of kCsrThreads, sized for >= ~3 CTAs per SM where the problem allows
static int chamfer_grad_csr_rows(int b, int n, int m) {
  static const int forced = [] {
    const char *e = getenv("MVP_CHAMFER_BWD_ROWS");
    return e ? atoi(e) : 0;
  }();
  if (forced >= kCsrThreads && forced % kCsrThreads == 0) return forced;
  int R = 4096;  // two CTAs of 512 threads are resident per SM: halve R until the grid fills one such wave
  while (R > kCsrThreads && (long long)b * (((n + R - 1) / R) + ((m + R - 1) / R)) < (long long)kNumSMs) R >>= 1;
  return R;
}

static bool chamfer_grad_csr_supported(int b, int n, int m) {
  return b <= 65535 && n <= kCsrMaxPts && m <= kCsrMaxPts;
}

static int chamfer_grad_csr_launch(int b, int n, int m, const float *xyz1, const float *xyz2, const float *gd1,
                                   const float *gd2, const int *idx1, const int *idx2, float *gx1, float *gx2,
                                   cudaStream_t s) {
  const int R = chamfer_grad_csr_rows(b, n, m);
  const size_t smem = sizeof(int) * (size_t)(R + 1) + sizeof(unsigned short) * (size_t)std::max(n, m);
  static size_t granted[kMaxDevices];
  const int rc = grant_dyn_smem(chamfer_grad_csr_kernel, smem, granted);
  if (rc) return rc;
  const int parts = (std::max(n, m) + R - 1) / R;
  chamfer_grad_csr_kernel<<<dim3(parts, 2, b), kCsrThreads, smem, s>>>(n, m, R, xyz1, xyz2, gd1, gd2, idx1, idx2, gx1,
                                                                      gx2);
  count_launch();
  return launch_status();
}

}  // namespace mvp

using namespace mvp;

namespace mvp {
// chamfer_fused.cu
bool chamfer_fused_supported(int b, int n, int m);
size_t chamfer_fused_workspace_bytes(int b, int n, int m);
int chamfer_fused_launch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1,
                         float *dist2, int *idx1, int *idx2, void *ws, size_t ws_bytes, cudaStream_t s);
// chamfer_grid.cu
bool chamfer_grid_supported(int b, int n, int m);
size_t chamfer_grid_workspace_bytes(int b, int n, int m);
int chamfer_grid_launch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1,
                        float *dist2, int *idx1, int *idx2, void *ws, size_t ws_bytes, cudaStream_t s);
}  // namespace mvp

MVP_API size_t mvp_chamfer_forward_workspace_bytes(int b, int n, int m) {
  if (b <= 0 || n <= 0 || m <= 0) return 16;
  size_t need = 16;  // one size serves every algorithm, so the caller need not know which one runs
  if (chamfer_fused_supported(b, n, m)) need = std::max(need, chamfer_fused_workspace_bytes(b, n, m));
  if (chamfer_grid_supported(b, n, m)) need = std::max(need, chamfer_grid_workspace_bytes(b, n, m));
  return need;
}

MVP_API int mvp_chamfer_forward_algo(int algo, int b, int n, int m, const float *xyz1, const float *xyz2,
                                     float *dist1, float *dist2, int *idx1, int *idx2, void *workspace,
                                     size_t workspace_bytes, mvp_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (algo < MVP_CHAMFER_AUTO || algo > MVP_CHAMFER_GRID) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || (n == 0 && m == 0)) return MVP_OK;
  if (n == 0 || m == 0) return MVP_ERR_INVALID_ARGUMENT;  // the reference reads out of bounds here
  if (!xyz1 || !xyz2 || !dist1 || !dist2 || !idx1 || !idx2) return MVP_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  const bool grid_ok = chamfer_grid_supported(b, n, m);
  if (algo == MVP_CHAMFER_GRID && !grid_ok) return MVP_ERR_INVALID_ARGUMENT;
  if (grid_ok && algo != MVP_CHAMFER_BRUTE) {
    if (!workspace || workspace_bytes < chamfer_grid_workspace_bytes(b, n, m)) return MVP_ERR_WORKSPACE;
    return chamfer_grid_launch(b, n, m, xyz1, xyz2, dist1, dist2, idx1, idx2, workspace, workspace_bytes, s);
  }
  if (chamfer_fused_supported(b, n, m)) {
    if (!workspace || workspace_bytes < chamfer_fused_workspace_bytes(b, n, m)) return MVP_ERR_WORKSPACE;
    return chamfer_fused_launch(b, n, m, xyz1, xyz2, dist1, dist2, idx1, idx2, workspace,
                                workspace_bytes, s);
  }
  int rc = chamfer_dir_launch(b, n, m, xyz1, xyz2, dist1, idx1, s);
  if (rc) return rc;
  return chamfer_dir_launch(b, m, n, xyz2, xyz1, dist2, idx2, s);
}

MVP_API int mvp_chamfer_forward(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1,
                                float *dist2, int *idx1, int *idx2, void *workspace,
                                size_t workspace_bytes, mvp_stream_t stream) {
  return mvp_chamfer_forward_algo(MVP_CHAMFER_AUTO, b, n, m, xyz1, xyz2, dist1, dist2, idx1, idx2, workspace,
                                  workspace_bytes, stream);
}

MVP_API int mvp_chamfer_backward_algo(int algo, int b, int n, int m, const float *xyz1, const float *xyz2,
                                      const float *graddist1, const float *graddist2, const int *idx1,
                                      const int *idx2, float *gradxyz1, float *gradxyz2, mvp_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (algo < MVP_CHAMFER_BWD_AUTO || algo > MVP_CHAMFER_BWD_SUMMED) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || (n == 0 && m == 0)) return MVP_OK;
  if (n == 0 || m == 0) return MVP_ERR_INVALID_ARGUMENT;
  if (!xyz1 || !xyz2 || !graddist1 || !graddist2 || !idx1 || !idx2 || !gradxyz1 || !gradxyz2)
    return MVP_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  if (algo == MVP_CHAMFER_BWD_SUMMED) {
    if (!chamfer_grad_csr_supported(b, n, m)) return MVP_ERR_INVALID_ARGUMENT;
    return chamfer_grad_csr_launch(b, n, m, xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1, gradxyz2, s);
  }
  const long long total = (long long)b * (n + m);
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
  auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  // four points per thread need enough points to fill the machine: below ~512k the one-point kernels are faster
  if (total >= (1 << 19) && n % 4 == 0 && m % 4 == 0 && al16(xyz1) && al16(xyz2) && al16(graddist1) && al16(graddist2) && al16(idx1) &&
      al16(idx2) && al16(gradxyz1) && al16(gradxyz2)) {
    const int grid4 = (int)std::min<long long>((total / 4 + 255) / 256, (long long)kNumSMs * 16);
    // One pass over zeroed gradients (31.7 us at the headline size) against two passes without a memset (33.8 us):
    // the second random gather of the neighbours' coordinates costs more than the memset and the extra coalesced
    // reductions.  MVP_CHAMFER_BWD_ONEPASS=0 selects the two-pass kernels.
    static const int one_pass = [] {
      const char *e = getenv("MVP_CHAMFER_BWD_ONEPASS");
      return e ? atoi(e) : 1;
    }();
    if (one_pass) {
      const size_t bytes1 = sizeof(float) * 3 * (size_t)b * n, bytes2 = sizeof(float) * 3 * (size_t)b * m;
      cudaError_t e1, e2 = cudaSuccess;
      if (reinterpret_cast<char *>(gradxyz1) + bytes1 == reinterpret_cast<char *>(gradxyz2)) {
        e1 = cudaMemsetAsync(gradxyz1, 0, bytes1 + bytes2, s);  // (the Python layer allocates both in one buffer)
      } else {
        e1 = cudaMemsetAsync(gradxyz1, 0, bytes1, s);
        e2 = cudaMemsetAsync(gradxyz2, 0, bytes2, s);
      }
      if (e1 != cudaSuccess || e2 != cudaSuccess) return (int)(e1 != cudaSuccess ? e1 : e2);
      chamfer_grad4_kernel<2><<<grid4, 256, 0, s>>>(b, n, m, xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1,
                                                    gradxyz2);
      count_launch();
      return launch_status();
    } else {
      chamfer_grad4_kernel<0><<<grid4, 256, 0, s>>>(b, n, m, xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1,
                                                    gradxyz2);
      chamfer_grad4_kernel<1><<<grid4, 256, 0, s>>>(b, n, m, xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1,
                                                    gradxyz2);
    }
  } else {
    chamfer_grad_kernel<false><<<grid, 256, 0, s>>>(b, n, m, xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1,
                                                    gradxyz2);
    chamfer_grad_kernel<true><<<grid, 256, 0, s>>>(b, n, m, xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1,
                                                   gradxyz2);
  }
  count_launch(2);
  return launch_status();
}

MVP_API int mvp_chamfer_backward(int b, int n, int m, const float *xyz1, const float *xyz2,
                                 const float *graddist1, const float *graddist2, const int *idx1,
                                 const int *idx2, float *gradxyz1, float *gradxyz2, mvp_stream_t stream) {
  return mvp_chamfer_backward_algo(MVP_CHAMFER_BWD_AUTO, b, n, m, xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1,
                                   gradxyz2, stream);
}
