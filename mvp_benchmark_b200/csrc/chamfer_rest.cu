// Chamfer distance: completion of the queries the grid search (chamfer_grid.cu) did not finish — far outside the
// other cloud (a prediction collapsed to a blob inside the ground truth's cube: PCN at initialisation, BASELINE config
// C2; disjoint clouds), inside or next to a very dense cell, against a degenerate grid, with non-finite coordinates.
// Exact like the brute-force kernels (same contraction, lowest index among equal minima; reference:
// utils/metrics/CD/chamfer3D/chamfer3D.cu:22-129), but the work is what the geometry requires, not n * m.
//
// The structure is a bounding-box hierarchy laid over the array the grid build already sorted by cell:
//   leaf     32 consecutive points of the sorted array (a piece of a row of cells: compact in y and z)
//   level 1  32 consecutive leaves (1024 points)          level 2  32 consecutive level-1 nodes (32768 points)
// Boxes are the exact minima / maxima of the coordinates, so the distance from a query to a box, evaluated with the
// SAME contraction as a candidate (gap_y^2 rounded, two fused multiply-adds), is a lower bound of every candidate's
// COMPUTED distance inside it without any slack: subtraction, multiplication and fma are monotone under rounding.
// Every node also carries the lowest original index below it: a node is skipped when its bound is above the best
// distance found, or equal to it with no index below the best one (the reference's tie rule, chamfer3D.cu:36,126).
// chamfer_rest_prep_kernel builds the boxes — only for clouds that some query still needs — and the list of work
// items; when the grid search finished everything (the normal case) both kernels leave at once.
// A cloud made of ONE point n times (hdr.pad[0]) is answered directly.  A non-finite query ends as (+inf, 0), the
// result of a strict `<` scan over distances that are never smaller than +inf.
#include <cstdlib>

#include "common.cuh"
#include "grid.cuh"
#include "sm100.cuh"

namespace mvp {

__device__ __forceinline__ uint32_t box_lb(const float4 lo, const float4 hi, float qx, float qy, float qz) {
  const float gx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.f);
  const float gy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.f);
  const float gz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.f);
  return __float_as_uint(sqdist(gx, gy, gz));
}

// box of the lanes' (lo, hi) and the lowest of their indices
__device__ __forceinline__ void warp_box(float (&lo)[3], float (&hi)[3], int &mi) {
#pragma unroll
  for (int off = 16; off; off >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
    }
  }
  mi = (int)redux_min_u32((uint32_t)mi);
}

// bits of v (< 2^16) spread to positions 0, 3, 6, ...
__device__ __forceinline__ u64 spread3(u64 v) {
  v = (v | v << 32) & 0x1f00000000ffffULL;
  v = (v | v << 16) & 0x1f0000ff0000ffULL;
  v = (v | v << 8) & 0x100f00f00f00f00fULL;
  v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
  v = (v | v << 2) & 0x1249249249249249ULL;
  return v;
}

// ------------------------------------------------------------------------------------------------ prep
// One CTA per (cloud, target side): nothing to do unless the list of the OTHER side's queries is non-empty.
//  1. does the target's grid hold a very dense cell (more than kDenseCell points: a tight cluster, or a unit cloud
//     next to far outliers)?  Such a cell is stored in arrival order: its leaves would all have the cell's box.
//  2. work items: the list in chunks of 32 entries, appended to plan list A (lanes over queries) or, for a dense
//     target, to plan list B (a warp per query: the queries inside a dense cell are not neighbours).
//  3. dense target of up to kMortonPts points: re-sorted along a Z-order curve (48-bit Morton code of the position in
//     the cloud's bounding cube, 16 bits per axis; bitonic sort of (code << 14 | original index) in shared memory),
//     the sorted array rewritten from the caller's coordinates: runs of the curve are compact in all three axes at
//     every scale.  The grid search is over by now: nothing else reads the cell order any more.
//     Otherwise the QUERY list is reordered instead (counting sort by the Z-order of blocks of cells of the queries'
//     own grid): the list arrives in the order of rows of cells — 32 consecutive entries are a stick across the
//     cloud — and leaves as compact groups that need the same nodes.
//  4. boxes: leaves (32 consecutive points), level 1 (32 leaves), level 2 (32 level-1 nodes); lo.w = the lowest
//     original index below the node.
constexpr int kPrepThreads = 1024;
constexpr int kMortonPts = 16384;
constexpr int kDenseCell = 512;
constexpr int kZBins = 4096;  // 16 x 16 x 16 blocks of cells
constexpr size_t kPrepSmem = sizeof(int) * kZBins + 2 * sizeof(int) * kMortonPts;  // >= sizeof(u64) * kMortonPts

__global__ void __launch_bounds__(kPrepThreads, 1)
chamfer_rest_prep_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2, GridWs W,
                         int plan_cap) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ int s_base, s_warp[32];
  pdl_wait();     // the query kernel's lists
  pdl_trigger();
  const int cloud = blockIdx.x, ts = blockIdx.y, qs = 1 - ts;
  const int li = qs * b + cloud;
  const int cnt = W.count[li];
  if (cnt == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const GridHdr h = W.hdr[ts * b + cloud];
  const int nt = ts ? m : n, nq = ts ? n : m;
  // ---- 1. the fullest cell
  int maxc = 0;
  if (h.valid) {
    const int *start = W.start[ts] + (size_t)cloud * (W.cap[ts] + 1);
    for (int c = tid; c < h.ncell; c += kPrepThreads) maxc = max(maxc, start[c + 1] - start[c]);
  }
  maxc = (int)redux_max_u32((uint32_t)maxc);
  if (lane == 0) s_warp[warp] = maxc;
  __syncthreads();
  maxc = (int)redux_max_u32((uint32_t)s_warp[lane]);
  const bool dense = maxc > kDenseCell;
  // ---- 2. work items
  const int nch = (cnt + 31) >> 5;
  if (tid == 0) s_base = atomicAdd(W.plan + (dense ? kPlanTotalB : kPlanTotal), nch);
  __syncthreads();
  for (int c = tid; c < nch; c += kPrepThreads)
    W.plan[kPlanItems + (dense ? plan_cap : 0) + s_base + c] = (int)(((unsigned)li << 15) | (unsigned)c);
  if (h.pad[0]) return;  // one point, n times: answered without boxes

  float4 *T = W.sorted[ts] + (size_t)cloud * nt;
  if (dense && nt <= kMortonPts) {
    // ---- 3a. the target along a Z-order curve
    u64 *s_key = reinterpret_cast<u64 *>(s_dyn);
    const float *P = (ts ? xyz2 : xyz1) + (size_t)cloud * nt * 3;
    int N = 64;
    while (N < nt) N <<= 1;
    const float ext = fmaxf(fmaxf((float)h.g[0], (float)h.g[1]), (float)h.g[2]) * h.s;  // >= the cloud's extent
    const float scale = 65536.f / ext;
    for (int i = tid; i < N; i += kPrepThreads) {
      u64 key = ~0ULL;
      if (i < nt) {
        const float4 c = T[i];
        const float ux = fminf(fmaxf((c.x - h.lo[0]) * scale, 0.f), 65535.f);
        const float uy = fminf(fmaxf((c.y - h.lo[1]) * scale, 0.f), 65535.f);
        const float uz = fminf(fmaxf((c.z - h.lo[2]) * scale, 0.f), 65535.f);
        const u64 code = spread3((u64)(unsigned)ux) | spread3((u64)(unsigned)uy) << 1 | spread3((u64)(unsigned)uz) << 2;
        key = code << 14 | (u64)(unsigned)__float_as_int(c.w);
      }
      s_key[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < (N >> 1); t += kPrepThreads) {
          const int i = 2 * t - (t & (j - 1)), l = i + j;  // the pair (i, i + j), bit j of i clear
          const u64 a = s_key[i], c = s_key[l];
          if ((a > c) == ((i & k) == 0)) s_key[i] = c, s_key[l] = a;
        }
        __syncthreads();
      }
    }
    for (int r = tid; r < nt; r += kPrepThreads) {
      const int o = (int)(s_key[r] & 16383ULL);
      T[r] = make_float4(__ldg(P + (size_t)o * 3), __ldg(P + (size_t)o * 3 + 1), __ldg(P + (size_t)o * 3 + 2), __int_as_float(o));
    }
    __syncthreads();
  } else if (!dense && cnt <= kMortonPts && cnt > 32) {
    // ---- 3b. the query list in compact groups
    const GridHdr hq = W.hdr[qs * b + cloud];
    if (hq.valid) {
      int *hist = reinterpret_cast<int *>(s_dyn);
      unsigned *code = reinterpret_cast<unsigned *>(hist + kZBins);
      int *og = reinterpret_cast<int *>(code + kMortonPts);
      const float *Pq = (qs ? xyz2 : xyz1) + (size_t)cloud * nq * 3;
      int *list = W.list[qs] + (size_t)cloud * nq;
      int sh = 0;
      while ((max(max(hq.g[0], hq.g[1]), hq.g[2]) >> sh) > 16) sh++;
      for (int c = tid; c < kZBins; c += kPrepThreads) hist[c] = 0;
      __syncthreads();
      auto spread4 = [](int v) { return (v & 1) | (v & 2) << 2 | (v & 4) << 4 | (v & 8) << 6; };
      for (int i = tid; i < cnt; i += kPrepThreads) {
        const int o = list[i];
        const int bx = cell_coord((__ldg(Pq + (size_t)o * 3) - hq.lo[0]) * hq.inv_s, hq.g[0]) >> sh;
        const int by = cell_coord((__ldg(Pq + (size_t)o * 3 + 1) - hq.lo[1]) * hq.inv_s, hq.g[1]) >> sh;
        const int bz = cell_coord((__ldg(Pq + (size_t)o * 3 + 2) - hq.lo[2]) * hq.inv_s, hq.g[2]) >> sh;
        const int key = spread4(bx) | spread4(by) << 1 | spread4(bz) << 2;
        code[i] = (unsigned)key << 16 | (unsigned)atomicAdd(&hist[key], 1);
        og[i] = o;
      }
      __syncthreads();
      // exclusive scan of the 4096 bins: four per thread
      int v[4], sum = 0;
#pragma unroll
      for (int e = 0; e < 4; e++) v[e] = hist[4 * tid + e], sum += v[e];
      int incl = sum;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += u;
      }
      __syncthreads();  // (s_warp was read above)
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        int u = s_warp[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int w = __shfl_up_sync(0xffffffffu, u, off);
          if (lane >= off) u += w;
        }
        s_warp[lane] = u;
      }
      __syncthreads();
      int run = incl - sum + (warp ? s_warp[warp - 1] : 0);
#pragma unroll
      for (int e = 0; e < 4; e++) hist[4 * tid + e] = run, run += v[e];
      __syncthreads();
      for (int i = tid; i < cnt; i += kPrepThreads) list[hist[code[i] >> 16] + (int)(code[i] & 0xffffu)] = og[i];
    }
  }

  // ---- 4. boxes
  const int nleaf = W.nleaf[ts], nl1 = (nleaf + 31) >> 5, nl2 = (nl1 + 31) >> 5;
  float4 *B0 = W.box[ts] + (size_t)cloud * 2 * (nleaf + nl1 + nl2), *B1 = B0 + 2 * nleaf, *B2 = B1 + 2 * nl1;
  const float inf = __int_as_float(0x7f800000);
  for (int leaf0 = warp; leaf0 < nleaf; leaf0 += 4 * (kPrepThreads / 32)) {
    float4 c[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {  // four leaves in flight
      const int pos = (leaf0 + u * (kPrepThreads / 32)) * 32 + lane;
      if (pos < nt) c[u] = T[pos];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int leaf = leaf0 + u * (kPrepThreads / 32), pos = leaf * 32 + lane;
      if (leaf < nleaf) {
        float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
        int mi = 0x7fffffff;
        if (pos < nt) {
          lo[0] = hi[0] = c[u].x, lo[1] = hi[1] = c[u].y, lo[2] = hi[2] = c[u].z;
          mi = __float_as_int(c[u].w);
        }
        warp_box(lo, hi, mi);  // fminf / fmaxf drop NaN coordinates: such a point can never win anyway
        if (lane == 0) {
          B0[2 * leaf] = make_float4(lo[0], lo[1], lo[2], __int_as_float(mi));
          B0[2 * leaf + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
        }
      }
    }
  }
  __syncthreads();
  for (int lvl = 0; lvl < 2; lvl++) {
    const float4 *src = lvl ? B1 : B0;
    float4 *dst = lvl ? B2 : B1;
    const int nsrc = lvl ? nl1 : nleaf, ndst = lvl ? nl2 : nl1;
    for (int j = warp; j < ndst; j += kPrepThreads / 32) {
      const int i = j * 32 + lane;
      float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
      int mi = 0x7fffffff;
      if (i < nsrc) {
        const float4 a = src[2 * i], c = src[2 * i + 1];
        lo[0] = a.x, lo[1] = a.y, lo[2] = a.z, hi[0] = c.x, hi[1] = c.y, hi[2] = c.z;
        mi = __float_as_int(a.w);
      }
      warp_box(lo, hi, mi);
      if (lane == 0) {
        dst[2 * j] = make_float4(lo[0], lo[1], lo[2], __int_as_float(mi));
        dst[2 * j + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ traversal
// A WARP owns 32 consecutive entries of a (direction, cloud) left-over list; every lane keeps ITS query and its own
// best (distance, index).  The chunk first meets the leaf nearest to its first unbounded query — every lane against
// the leaf's 32 points — which bounds all lanes within a few cells of their answers.  Then one of two traversals:
//
//   lanes over queries   (the chunk's queries need the same few nodes: disjoint clouds, a far compact target)
//       The children of a node are looked at one after the other, nearest to the chunk's bounding box first (one
//       box-against-box bound per child, lanes over the 32 children; redux.sync.min picks the next one and stops at the
//       loosest best of the chunk).  A child is descended into when ANY lane's own bound reaches it (one ballot), and
//       a leaf's 32 points cost 32 x (one broadcast load + the contraction + a two-word compare) for 32 queries at
//       once — no cross-lane reduction anywhere.
//   a warp per query     (the queries need different nodes: a target much denser than the spacing of the queries,
//       queries inside a dense cell — list B of the plan)
//       The queries are taken one at a time, LANES OVER THE 32 CHILDREN of the node it is looking at: one box test
//       (or one candidate) per lane, the nearest live child by redux.sync.min over the bit patterns of the bounds
//       (non-negative floats order like unsigned integers) + one ballot.  Each query also starts from the point that
//       won for the previous one.
// The choice is made per chunk (see "which traversal?" below), and a chunk that has opened kBatchLeaves leaves lanes
// over queries finishes query by query with the bounds it has by then.
constexpr int kRestWarps = 4;
constexpr int kRestThreads = 32 * kRestWarps;
constexpr int kRestCtas = kNumSMs * 10;  // one resident wave; the warps draw work items from a ticket counter
constexpr float kCoherentRatio = 4.f;
constexpr int kBatchLeaves = 12;

// can a child with bound lb and lowest index mi improve on (best, bidx)?
__device__ __forceinline__ bool reaches(uint32_t lb, int mi, uint32_t best, int bidx) {
  return lb < best || (lb == best && mi < bidx);
}

__global__ void __launch_bounds__(kRestThreads, 10)
chamfer_rest_kernel(int b, int n, int m, GridWs W, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                    float *__restrict__ dist1, float *__restrict__ dist2, int *__restrict__ idx1, int *__restrict__ idx2,
                    int plan_cap) {
  pdl_wait();     // the plan written by chamfer_rest_prep_kernel
  const int total_a = W.plan[kPlanTotal], total = total_a + W.plan[kPlanTotalB];
  if (total == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  const uint32_t kInf = 0x7f800000u, kNone = 0xffffffffu;
  const int nwarps = gridDim.x * kRestWarps;
  const float finf = __uint_as_float(kInf);

  int item = blockIdx.x * kRestWarps + warp;  // the first item is static, the following ones are drawn
  while (item < total) {
    const bool list_b = item >= total_a;
    const unsigned it = (unsigned)W.plan[kPlanItems + (list_b ? plan_cap + item - total_a : item)];
    const int li = (int)(it >> 15), c0 = (int)(it & 32767u) * 32;
    const int dir = li >= b ? 1 : 0, cloud = dir ? li - b : li;
    const int cnt = W.count[li];
    const int nq = dir ? m : n, nt = dir ? n : m, ts = 1 - dir;
    const float4 *T = W.sorted[ts] + (size_t)cloud * nt;
    const float *Pq = (dir ? xyz2 : xyz1) + (size_t)cloud * nq * 3;
    const int *list = W.list[dir] + (size_t)cloud * nq;
    float *dist = (dir ? dist2 : dist1) + (size_t)cloud * nq;
    int *idx = (dir ? idx2 : idx1) + (size_t)cloud * nq;

    const bool present = c0 + lane < cnt;
    int orig = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    uint32_t best = kInf;
    int bidx = 0x7fffffff;
    if (present) {
      orig = list[c0 + lane];
      qx = __ldg(Pq + (size_t)orig * 3), qy = __ldg(Pq + (size_t)orig * 3 + 1), qz = __ldg(Pq + (size_t)orig * 3 + 2);
      best = __float_as_uint(dist[orig]);  // what the grid search found before it gave up (+inf: nothing)
      bidx = idx[orig];
    }
    if (!(best < kInf) || bidx < 0 || bidx >= nt) best = kInf, bidx = 0x7fffffff;  // no usable bound
    const bool finite = fabsf(qx) + fabsf(qy) + fabsf(qz) < 3.0e38f;  // false for NaN / inf
    const bool act = present && finite;
    if (W.hdr[ts * b + cloud].pad[0]) {
      // the target cloud is ONE point, n times (a collapsed prediction): its lowest index is the answer, at the
      // distance the reference's scan would find first — or (+inf, 0) when that distance is not below +inf
      if (present) {
        const float4 c = __ldg(T);
        const float d = sqdist(c.x - qx, c.y - qy, c.z - qz);
        dist[orig] = d < finf ? d : finf;
        idx[orig] = 0;
      }
    } else {
      const int nleaf = W.nleaf[ts], nl1 = (nleaf + 31) >> 5, nl2 = (nl1 + 31) >> 5;
      const float4 *B0 = W.box[ts] + (size_t)cloud * 2 * (nleaf + nl1 + nl2), *B1 = B0 + 2 * nleaf, *B2 = B1 + 2 * nl1;
      if (!act) best = 0u, bidx = -1;  // nothing reaches such a lane, nothing improves it

      // the points of a leaf against every lane's query
      auto eval_leaf = [&](int leaf) {
        const float4 *p = T + (size_t)leaf * 32;
        const int np = min(32, nt - leaf * 32);
#pragma unroll 4
        for (int c = 0; c < np; c++) {
          const float4 v = __ldg(p + c);
          const uint32_t d = __float_as_uint(sqdist(v.x - qx, v.y - qy, v.z - qz));  // >= +0, or NaN (above +inf as an integer)
          const int vi = __float_as_int(v.w);
          if (d < best || (d == best && vi < bidx)) best = d, bidx = vi;
        }
      };
      // the bounding box of the chunk's queries, for the box-against-box bounds
      float cl[3], ch[3];
      {
        cl[0] = act ? qx : finf, cl[1] = act ? qy : finf, cl[2] = act ? qz : finf;
        ch[0] = act ? qx : -finf, ch[1] = act ? qy : -finf, ch[2] = act ? qz : -finf;
#pragma unroll
        for (int off = 16; off; off >>= 1) {
#pragma unroll
          for (int a = 0; a < 3; a++) {
            cl[a] = fminf(cl[a], __shfl_xor_sync(full, cl[a], off));
            ch[a] = fmaxf(ch[a], __shfl_xor_sync(full, ch[a], off));
          }
        }
      }
      // lanes over the children [first, first + 32) of a level (boxes at B): the bound between the chunk's box and the
      // child's — below every lane's own bound for that child
      auto coarse = [&](const float4 *B, int first, int count) -> uint32_t {
        const int i = min(first + lane, count - 1);
        const float4 lo = B[2 * i], hi = B[2 * i + 1];
        const float gx = fmaxf(fmaxf(lo.x - ch[0], cl[0] - hi.x), 0.f);
        const float gy = fmaxf(fmaxf(lo.y - ch[1], cl[1] - hi.y), 0.f);
        const float gz = fmaxf(fmaxf(lo.z - ch[2], cl[2] - hi.z), 0.f);
        return first + lane < count ? __float_as_uint(sqdist(gx, gy, gz)) : kNone;
      };
      // the nearest child not yet taken whose bound is within the chunk's loosest best, or -1
      auto next_child = [&](uint32_t &lb) -> int {
        const uint32_t mn = redux_min_u32(lb);
        if (mn == kNone || mn > redux_max_u32(best)) return -1;
        const int j = __ffs(__ballot_sync(full, lb == mn)) - 1;
        if (lane == j) lb = kNone;
        return j;
      };

      // ---- first bound: the leaf nearest to the first unbounded query (lanes over children on the way down)
      int seed_leaf = -1;
      const unsigned unb = __ballot_sync(full, act && best == kInf);
      if (unb) {
        const int l0 = __ffs(unb) - 1;
        const float sx = __shfl_sync(full, qx, l0), sy = __shfl_sync(full, qy, l0), sz = __shfl_sync(full, qz, l0);
        int node = 0;
        if (nl2 > 1) {
          const uint32_t lb = lane < nl2 ? box_lb(B2[2 * lane], B2[2 * lane + 1], sx, sy, sz) : kNone;
          node = __ffs(__ballot_sync(full, lb == redux_min_u32(lb))) - 1;
        }
        {
          const int i1 = node * 32 + lane;
          const uint32_t lb = i1 < nl1 ? box_lb(B1[2 * i1], B1[2 * i1 + 1], sx, sy, sz) : kNone;
          node = node * 32 + __ffs(__ballot_sync(full, lb == redux_min_u32(lb))) - 1;
        }
        {
          const int lf = node * 32 + lane;
          const uint32_t lb = lf < nleaf ? box_lb(B0[2 * lf], B0[2 * lf + 1], sx, sy, sz) : kNone;
          seed_leaf = node * 32 + __ffs(__ballot_sync(full, lb == redux_min_u32(lb))) - 1;
        }
        eval_leaf(seed_leaf);
      }
      // ---- which traversal?  Lanes over queries pays when the chunk is small against the target's leaves (each leaf
      // it opens then serves many lanes): the chunk's box against the seed leaf's, by their squared diagonals.  A chunk
      // without a seed (every query arrived with a bound: it gave up inside a crowded neighbourhood) goes query by query.
      bool batch = !list_b && seed_leaf >= 0;
      if (batch) {
        const float4 lo = B0[2 * seed_leaf], hi = B0[2 * seed_leaf + 1];
        const float dl = sqdist(hi.x - lo.x, hi.y - lo.y, hi.z - lo.z), dc = sqdist(ch[0] - cl[0], ch[1] - cl[1], ch[2] - cl[2]);
        batch = dc <= kCoherentRatio * dl;
      }
      int opened = 0;  // leaves opened lanes-over-queries; past kBatchLeaves the chunk finishes query by query
      if (batch) {
        for (int j2 = 0; j2 < nl2 && opened <= kBatchLeaves; j2++) {
          if (nl2 > 1) {
            const float4 lo = B2[2 * j2], hi = B2[2 * j2 + 1];
            if (!__any_sync(full, reaches(box_lb(lo, hi, qx, qy, qz), __float_as_int(lo.w), best, bidx))) continue;
          }
          uint32_t c1 = coarse(B1, j2 * 32, nl1);
          for (int k1; opened <= kBatchLeaves && (k1 = next_child(c1)) >= 0;) {
            const int j1 = j2 * 32 + k1;
            {
              const float4 lo = B1[2 * j1], hi = B1[2 * j1 + 1];
              if (!__any_sync(full, reaches(box_lb(lo, hi, qx, qy, qz), __float_as_int(lo.w), best, bidx))) continue;
            }
            uint32_t cf = coarse(B0, j1 * 32, nleaf);
            for (int kl; opened <= kBatchLeaves && (kl = next_child(cf)) >= 0;) {
              const int leaf = j1 * 32 + kl;
              if (leaf == seed_leaf) continue;
              const float4 lo = B0[2 * leaf], hi = B0[2 * leaf + 1];
              if (!__any_sync(full, reaches(box_lb(lo, hi, qx, qy, qz), __float_as_int(lo.w), best, bidx))) continue;
              eval_leaf(leaf);
              opened++;
            }
          }
        }
      }
      if (!batch || opened > kBatchLeaves) {
        // a cloud of up to 32768 points has ONE level-2 node: every lane keeps its level-1 box for the whole chunk
        float4 lo1c = make_float4(0.f, 0.f, 0.f, 0.f), hi1c = lo1c;
        if (nl2 == 1 && lane < nl1) lo1c = B1[2 * lane], hi1c = B1[2 * lane + 1];
        float wx = 0.f, wy = 0.f, wz = 0.f;  // the last point that won for a query of this chunk: the next query's first candidate
        int widx = -1;
        unsigned todo = __ballot_sync(full, act);
        while (todo) {
          const int k = __ffs(todo) - 1;
          todo &= todo - 1;
          const float kx = __shfl_sync(full, qx, k), ky = __shfl_sync(full, qy, k), kz = __shfl_sync(full, qz, k);
          uint32_t kb = __shfl_sync(full, best, k);
          int ki = __shfl_sync(full, bidx, k);
          if (widx >= 0) {
            const uint32_t d = __float_as_uint(sqdist(wx - kx, wy - ky, wz - kz));
            if (d < kb || (d == kb && widx < ki)) kb = d, ki = widx;
          }
          // children of a level-1 node: leaf boxes, then the leaves' points
          auto visit_l1 = [&](int node1) {
            const int leaf0 = node1 * 32, lf = min(leaf0 + lane, nleaf - 1);  // (a repeated child changes nothing)
            const float4 lol = B0[2 * lf], hil = B0[2 * lf + 1];
            uint32_t lbl = box_lb(lol, hil, kx, ky, kz);
            const int mil = __float_as_int(lol.w);
            while (true) {
              const uint32_t key = reaches(lbl, mil, kb, ki) ? lbl : kNone;
              const uint32_t ml = redux_min_u32(key);
              if (ml == kNone) break;
              const int jl = __ffs(__ballot_sync(full, key == ml)) - 1;
              if (lane == jl) lbl = kNone;
              // ---- a leaf: one candidate per lane (the last leaf of a cloud repeats its last point)
              const int pos = min((leaf0 + jl) * 32 + lane, nt - 1);
              const float4 c = __ldg(T + pos);
              const uint32_t db = __float_as_uint(sqdist(c.x - kx, c.y - ky, c.z - kz));
              const uint32_t dm = redux_min_u32(db);
              if (dm <= kb) {
                const int ci = db == dm ? __float_as_int(c.w) : 0x7fffffff;
                const int im = (int)redux_min_u32((uint32_t)ci);
                if (dm < kb || im < ki) {
                  kb = dm, ki = im;
                  const int wl = __ffs(__ballot_sync(full, ci == im)) - 1;
                  wx = __shfl_sync(full, c.x, wl), wy = __shfl_sync(full, c.y, wl), wz = __shfl_sync(full, c.z, wl);
                  widx = im;
                }
              }
            }
          };
          // children of a level-2 node: level-1 boxes (this lane's bound lb1; beyond the last one: kNone)
          auto visit_l2 = [&](int node2, uint32_t lb1, int mi1) {
            while (true) {
              const uint32_t key = reaches(lb1, mi1, kb, ki) ? lb1 : kNone;
              const uint32_t m1 = redux_min_u32(key);
              if (m1 == kNone) break;
              const int j1 = __ffs(__ballot_sync(full, key == m1)) - 1;
              if (lane == j1) lb1 = kNone;
              visit_l1(node2 * 32 + j1);
            }
          };
          if (nl2 == 1) {
            visit_l2(0, lane < nl1 ? box_lb(lo1c, hi1c, kx, ky, kz) : kNone, __float_as_int(lo1c.w));
          } else {
            const float4 lo2 = B2[2 * min(lane, nl2 - 1)], hi2 = B2[2 * min(lane, nl2 - 1) + 1];
            uint32_t lb2 = lane < nl2 ? box_lb(lo2, hi2, kx, ky, kz) : kNone;
            while (true) {
              const uint32_t key = reaches(lb2, __float_as_int(lo2.w), kb, ki) ? lb2 : kNone;
              const uint32_t m2 = redux_min_u32(key);
              if (m2 == kNone) break;
              const int j2 = __ffs(__ballot_sync(full, key == m2)) - 1;
              if (lane == j2) lb2 = kNone;
              const int i1 = j2 * 32 + lane, c1 = min(i1, nl1 - 1);
              const float4 lo1 = B1[2 * c1], hi1 = B1[2 * c1 + 1];
              visit_l2(j2, i1 < nl1 ? box_lb(lo1, hi1, kx, ky, kz) : kNone, __float_as_int(lo1.w));
            }
          }
          if (lane == k) best = kb, bidx = ki;
        }
      }
      if (present) {
        const bool found = act && best < kInf && bidx != 0x7fffffff;
        dist[orig] = found ? __uint_as_float(best) : finf;
        idx[orig] = found ? bidx : 0;
      }
    }
    int nxt = 0;
    if (lane == 0) nxt = nwarps + atomicAdd(W.plan + kPlanTicket2, 1);
    item = __shfl_sync(full, nxt, 0);
  }
}

int chamfer_rest_launch(int b, int n, int m, const GridWs &W, const float *xyz1, const float *xyz2, float *dist1,
                        float *dist2, int *idx1, int *idx2, cudaStream_t s) {
  static size_t granted[kMaxDevices];
  const int rc = grant_dyn_smem(chamfer_rest_prep_kernel, kPrepSmem, granted, 0);
  if (rc) return rc;
  const int plan_cap = rest_plan_cap(b, n, m);
  static const int pdl = [] {
    const char *e = getenv("MVP_PDL");
    return e ? atoi(e) : 1;
  }();
  if (pdl) {
    cudaError_t e = launch_pdl(chamfer_rest_prep_kernel, dim3(b, 2), dim3(kPrepThreads), kPrepSmem, s, b, n, m, xyz1, xyz2, W,
                               plan_cap);
    if (e == cudaSuccess)
      e = launch_pdl(chamfer_rest_kernel, dim3(kRestCtas), dim3(kRestThreads), 0, s, b, n, m, W, xyz1, xyz2, dist1, dist2, idx1,
                     idx2, plan_cap);
    if (e != cudaSuccess) return (int)e;
  } else {
    chamfer_rest_prep_kernel<<<dim3(b, 2), kPrepThreads, kPrepSmem, s>>>(b, n, m, xyz1, xyz2, W, plan_cap);
    chamfer_rest_kernel<<<kRestCtas, kRestThreads, 0, s>>>(b, n, m, W, xyz1, xyz2, dist1, dist2, idx1, idx2, plan_cap);
  }
  count_launch(2);
  return launch_status();
}

}  // namespace mvp
