// Chamfer distance: completion of the queries the grid search (chamfer_grid.cu) did not finish — far outside the
// other cloud (a prediction collapsed to a blob inside the ground truth's cube: PCN at initialisation, BASELINE config
// C2), inside or next to a very dense cell, against a degenerate grid, with non-finite coordinates.  Replaces the
// hand-over to the brute-force kernels: exact like them (same contraction, lowest index among equal minima,
// reference: utils/metrics/CD/chamfer3D/chamfer3D.cu:22-129), but the work is what the geometry requires, not n * m.
//
// A WARP owns 32 consecutive entries of a (direction, cloud) left-over list — appended by neighbouring lanes of the
// query kernel, i.e. neighbours in space — and alternates between two views of its lanes:
//   lanes over ROWS        the rows (y, z) of cells of the target grid inside the rectangle the 32 queries' bounds
//                          span, 32 at a time: a lane tests its row against every query (broadcast from shared
//                          memory) with the conservative lower bound of grid.cuh, in world units, and keeps the union
//                          of the x-ranges of cells the bounds leave.  A row is one contiguous range of the sorted
//                          target array; the surviving ranges of a batch form one candidate list (prefix scan).
//   lanes over CANDIDATES  the list is staged tile by tile in shared memory (structure of arrays) and taken into
//                          registers, two candidates per 64-bit register pair; the queries are broadcast one by one:
//                          packed fp32x2 distances (3 issue slots per 2 candidates), FMNMX3, one redux.sync.min over the
//                          bit patterns, one ballot naming the lanes that hold the minimum.  Which candidate of such
//                          a lane it was, and the lowest original index among equal minima, is settled afterwards by
//                          the query's own lane re-evaluating that lane's few slots — or on the spot when many lanes
//                          tie (coincident points).  Every tile tightens the bounds the next batch of rows is tested with.
// A query that arrives without a bound (nothing of the other cloud anywhere near) first meets 256 points spread evenly
// over the sorted target array, which bounds it within a few cells of its true distance.  A degenerate grid is one cell
// holding the whole cloud: the same loop with a single row.  A non-finite query ends as (+inf, 0), the result of the
// reference's strict `<` scan over distances that are never smaller than +inf.
#include <type_traits>

#include "common.cuh"
#include "grid.cuh"
#include "sm100.cuh"

namespace mvp {

constexpr int kDqTile = 256;  // candidate slots per shared-memory tile: 8 per lane

// A tile of candidates in shared memory, structure of arrays: slot s of the tile is x[s], y[s], z[s], id[s].  A lane
// owns the slots {2 lane + 64 h, 2 lane + 64 h + 1}, h < K2: each pair arrives as one 64-bit load, already in the
// register pair the packed instructions want.
struct DqTile {
  float x[kDqTile], y[kDqTile], z[kDqTile];
  int id[kDqTile];
};

// Candidate slots [0, navail) of `C` against the queries of the lanes in `qmask` (sQ: x, y, z per lane): per query the
// minimum over the tile as a bit pattern and `twin`: the mask of the (one or two) lanes that hold it, or — when more
// lanes tie — the complement of the lowest original index among the tied candidates, settled on the spot.
template <int K2>
__device__ __forceinline__ void dq_tile(const DqTile &C, const float4 *sQ, unsigned qmask, int lane, int navail,
                                        uint32_t &tbest, unsigned &twin) {
  const float nanv = __int_as_float(0x7fffffff);  // an absent slot: its distance is NaN, which a minimum ignores
  u64 X[K2], Y[K2], Z[K2];
#pragma unroll
  for (int h = 0; h < K2; h++) {
    const int s0 = 2 * lane + 64 * h;
    X[h] = *reinterpret_cast<const u64 *>(&C.x[s0]);
    Y[h] = *reinterpret_cast<const u64 *>(&C.y[s0]);
    Z[h] = *reinterpret_cast<const u64 *>(&C.z[s0]);
    if (s0 + 1 >= navail) {  // the tile's tail
      float xl, xh;
      unpack2(X[h], xl, xh);
      X[h] = pack2(s0 < navail ? xl : nanv, nanv);
    }
  }
#pragma unroll 1
  while (qmask) {
    const int l = __ffs(qmask) - 1;
    qmask &= qmask - 1;
    const float4 q = sQ[l];
    const u64 qx = pack2(q.x, q.x), qy = pack2(q.y, q.y), qz = pack2(q.z, q.z);
    float ml;
#pragma unroll
    for (int h = 0; h < K2; h++) {
      const u64 dx = sub2(X[h], qx), dy = sub2(Y[h], qy), dz = sub2(Z[h], qz);
      const u64 d = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
      float dl, dh;
      unpack2(d, dl, dh);
      ml = h == 0 ? fminf(dl, dh) : min3(ml, dl, dh);
    }
    const uint32_t mb = __float_as_uint(ml);  // d >= +0 or NaN: the bit patterns order like the values, NaN last
    const uint32_t mw = redux_min_u32(mb);
    unsigned bal = __ballot_sync(0xffffffffu, mb == mw);
    if (__popc(bal) > 2) {
      // many lanes hold the minimum (coincident points, lattices): settle the lowest original index here, lanes over
      // candidates, instead of letting the query's lane walk every tied lane's slots afterwards.  (warp-uniform branch)
      int ti = 0x7fffffff;
#pragma unroll
      for (int h = 0; h < K2; h++) {
        const u64 dx = sub2(X[h], qx), dy = sub2(Y[h], qy), dz = sub2(Z[h], qz);
        const u64 d = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
        float dl, dh;
        unpack2(d, dl, dh);
        const int s0 = 2 * lane + 64 * h;
        if (__float_as_uint(dl) == mw) ti = min(ti, C.id[s0]);
        if (__float_as_uint(dh) == mw) ti = min(ti, C.id[s0 + 1]);
      }
      bal = ~redux_min_u32((uint32_t)ti);  // the COMPLEMENT of the index (< 2^24): more than two bits set, unlike a mask kept below
    }
    const bool me = lane == l;
    tbest = me ? mw : tbest;
    twin = me ? bal : twin;
  }
}

// The exact winner among the slots of the lanes in `win` (lanes over queries): the lowest original index among the
// candidates whose distance has exactly the winning bit pattern.
__device__ __forceinline__ int dq_resolve(const DqTile &C, float qx, float qy, float qz, uint32_t tbest, unsigned win,
                                          int navail) {
  if (__popc(win) > 2) return (int)~win;  // settled by dq_tile
  int ti = 0x7fffffff;
  while (win) {
    const int wl = __ffs(win) - 1;
    win &= win - 1;
    for (int s0 = 2 * wl; s0 < navail; s0 += 64) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int s = s0 + e;
        if (s < navail) {
          const float d = sqdist(C.x[s] - qx, C.y[s] - qy, C.z[s] - qz);
          if (__float_as_uint(d) == tbest) ti = min(ti, C.id[s]);
        }
      }
    }
  }
  return ti;
}

// Stages the slots [t0, t0 + navail) of a candidate list made of NR (a power of two) contiguous ranges of the sorted
// target array: range r starts at list slot off[r] and at array position pos[r].  Slot s lies in the last range
// whose first slot is <= s (empty ranges share their successor's first slot and are never chosen).
template <int NR>
__device__ __forceinline__ void dq_stage(DqTile &C, const float4 *__restrict__ T, const int *off, const int *pos, int t0,
                                         int navail, int lane) {
#pragma unroll 1
  for (int k0 = 0; k0 < kDqTile / 32 && 32 * k0 < navail; k0 += 4) {  // four slots per lane in flight
    float4 cv[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int sl = lane + 32 * (k0 + k);
      if (sl < navail) {
        const int g = t0 + sl;
        int r = 0;
#pragma unroll
        for (int st = NR / 2; st; st >>= 1) r += off[r + st] <= g ? st : 0;
        cv[k] = __ldg(T + pos[r] + (g - off[r]));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int sl = lane + 32 * (k0 + k);
      if (sl < navail) C.x[sl] = cv[k].x, C.y[sl] = cv[k].y, C.z[sl] = cv[k].z, C.id[sl] = __float_as_int(cv[k].w);
    }
  }
}

template <typename F>
__device__ __forceinline__ void dq_dispatch(int navail, F &&f) {
  const int k2 = (navail + 63) >> 6;
  if (k2 == 1) f(std::integral_constant<int, 1>());
  else if (k2 == 2) f(std::integral_constant<int, 2>());
  else if (k2 == 3) f(std::integral_constant<int, 3>());
  else f(std::integral_constant<int, 4>());
}

constexpr int kDrWarps = 4;
constexpr int kDrThreads = 32 * kDrWarps;
constexpr int kDrCtasPerList = 16;  // grid.x: CTAs past the end of a list leave at once

struct DrSmem {
  DqTile tile[kDrWarps];
  float4 qry[kDrWarps][32];   // x, y, z, current bound (squared distance; -1: needs nothing)
  float4 cell[kDrWarps][32];  // the query in cell units of the target grid
  int off[kDrWarps][32], pos[kDrWarps][32];
};

__global__ void __launch_bounds__(kDrThreads, 6)
chamfer_rest_kernel(int b, int n, int m, GridWs W, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                    float *__restrict__ dist1, float *__restrict__ dist2, int *__restrict__ idx1, int *__restrict__ idx2) {
  __shared__ __align__(16) DrSmem S;
  const int li = blockIdx.y, dir = li >= b ? 1 : 0, cloud = dir ? li - b : li;
  const int cnt = __ldg(W.count + li);
  if ((int)(blockIdx.x * kDrThreads) >= cnt) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  DqTile &C = S.tile[warp];
  float4 *sQ = S.qry[warp], *sU = S.cell[warp];
  int *sOff = S.off[warp], *sPos = S.pos[warp];
  const int nq = dir ? m : n, nt = dir ? n : m, ts = 1 - dir;
  const GridHdr h = W.hdr[ts * b + cloud];
  const int gx = h.g[0], gy = h.g[1], gz = h.g[2];
  const int *start = (ts ? W.start[1] : W.start[0]) + (size_t)cloud * ((ts ? W.cap[1] : W.cap[0]) + 1);
  const float4 *T = (ts ? W.sorted[1] : W.sorted[0]) + (size_t)cloud * nt;
  const float *Pq = (dir ? xyz2 : xyz1) + (size_t)cloud * nq * 3;
  const int *list = (dir ? W.list[1] : W.list[0]) + (size_t)cloud * nq;
  float *dist = (dir ? dist2 : dist1) + (size_t)cloud * nq;
  int *idx = (dir ? idx2 : idx1) + (size_t)cloud * nq;
  const float inf = __int_as_float(0x7f800000), shr = 1.f - 1e-5f;

  for (int c0 = (blockIdx.x * kDrWarps + warp) * 32; c0 < cnt; c0 += gridDim.x * kDrThreads) {
    const bool present = c0 + lane < cnt;
    int orig = 0;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t best = 0x7f800000u;
    int bidx = 0x7fffffff;
    if (present) {
      orig = __ldg(list + c0 + lane);
      q.x = __ldg(Pq + (size_t)orig * 3), q.y = __ldg(Pq + (size_t)orig * 3 + 1), q.z = __ldg(Pq + (size_t)orig * 3 + 2);
      best = __float_as_uint(dist[orig]);  // what the grid search found before it gave up (+inf: nothing)
      bidx = idx[orig];
    }
    const float ux = (q.x - h.lo[0]) * h.inv_s, uy = (q.y - h.lo[1]) * h.inv_s, uz = (q.z - h.lo[2]) * h.inv_s;
    const bool finite = fabsf(q.x) + fabsf(q.y) + fabsf(q.z) < 3.0e38f && fabsf(ux) + fabsf(uy) + fabsf(uz) < 3.0e38f;
    const bool active = present && finite;
    if (!(best < 0x7f800000u) || bidx < 0 || bidx >= nt) best = 0x7f800000u, bidx = 0x7fffffff;  // no usable bound
    if (h.pad[0]) {
      // the target cloud is ONE point, n times (a collapsed prediction): its lowest index is the answer, at the
      // distance the reference's scan would find first — or (+inf, 0) when that distance is not below +inf
      if (present) {
        const float4 c = __ldg(T);
        const float d = sqdist(c.x - q.x, c.y - q.y, c.z - q.z);
        dist[orig] = d < inf ? d : inf;
        idx[orig] = 0;
      }
      continue;
    }
    const unsigned grp = __ballot_sync(0xffffffffu, active);
    __syncwarp();
    sQ[lane] = make_float4(q.x, q.y, q.z, active ? __uint_as_float(best) : -1.f);
    sU[lane] = make_float4(ux, uy, uz, 0.f);
    __syncwarp();
    if (grp) {
      const int l0 = __ffs(grp) - 1;
      auto evaluate = [&](int navail, unsigned qmask) {  // the staged tile against the queries in qmask; bounds tightened in place
        uint32_t tbest = 0x7fffffffu;
        unsigned twin = 0;
        dq_dispatch(navail, [&](auto k2) { dq_tile<decltype(k2)::value>(C, sQ, qmask, lane, navail, tbest, twin); });
        if (((qmask >> lane) & 1u) && tbest <= best) {
          const int ti = dq_resolve(C, q.x, q.y, q.z, tbest, twin, navail);
          if (tbest < best || ti < bidx) best = tbest, bidx = ti;
        }
      };
      if (__any_sync(0xffffffffu, active && best == 0x7f800000u)) {
        // A query without any bound would need every row.  Probe first: 256 points spread evenly over the sorted
        // target array, i.e. over the occupied cells, bound every query of the warp within a few cells of its
        // distance; then the 256 points around the best of them in the sorted order (its cell and the cells next to
        // it along x) — the warp's queries are neighbours, so are their nearest points — tighten that to about a cell.
        __syncwarp();
        for (int sl = lane; sl < kDqTile; sl += 32) {
          const float4 c = __ldg(T + (int)(((long long)sl * nt) / kDqTile));
          C.x[sl] = c.x, C.y[sl] = c.y, C.z[sl] = c.z, C.id[sl] = sl;  // (the slot, not the index: see below)
        }
        __syncwarp();
        {
          uint32_t tbest = 0x7fffffffu;
          unsigned twin = 0;
          dq_tile<4>(C, sQ, grp, lane, kDqTile, tbest, twin);
          const int slot = active ? dq_resolve(C, q.x, q.y, q.z, tbest, twin, kDqTile) : 0x7fffffff;
          if (active && slot < kDqTile && tbest <= best) {  // the lane's own nearest sample is a bound too
            const int ci = __float_as_int(__ldg(T + (int)(((long long)slot * nt) / kDqTile)).w);
            if (tbest < best || ci < bidx) best = tbest, bidx = ci;
          }
          // the sample nearest to the first active query names the neighbourhood the whole warp looks at next
          const int s0 = __shfl_sync(0xffffffffu, slot, l0);
          const int centre = s0 < kDqTile ? (int)(((long long)s0 * nt) / kDqTile) : 0;
          const int first = max(0, min(centre - kDqTile / 2, nt - kDqTile));
          const int navail = min(kDqTile, nt - first);
          __syncwarp();
          for (int sl = lane; sl < navail; sl += 32) {
            const float4 c = __ldg(T + first + sl);
            C.x[sl] = c.x, C.y[sl] = c.y, C.z[sl] = c.z, C.id[sl] = __float_as_int(c.w);
          }
          __syncwarp();
          evaluate(navail, grp);
        }
        __syncwarp();
        if (active) sQ[lane].w = __uint_as_float(best);
        __syncwarp();
      }
      // The rows are shared by queries that see the target from the same side: per axis, below the grid, inside its
      // extent, or above it (27 classes).  A row keeps ONE interval of cells — the union over the queries that need
      // it — and the union over queries on opposite sides of a compact cloud would be the whole row.
      const int cls = (ux < 0.f ? 0 : ux >= (float)gx ? 2 : 1) + 3 * (uy < 0.f ? 0 : uy >= (float)gy ? 2 : 1) +
                      9 * (uz < 0.f ? 0 : uz >= (float)gz ? 2 : 1);
      unsigned remaining = grp;
      while (remaining) {
      const int lead_cls = __shfl_sync(0xffffffffu, cls, __ffs(remaining) - 1);
      const unsigned sub = __ballot_sync(0xffffffffu, active && cls == lead_cls) & remaining;
      remaining &= ~sub;
      const bool mine = (sub >> lane) & 1u;
      // rectangle of rows any query's bound reaches (the whole grid for an infinite bound); one row for a degenerate grid
      int ylo = 0, yhi = 0, zlo = 0, zhi = 0;
      if (h.valid) {
        float fy0 = (float)gy, fy1 = -1.f, fz0 = (float)gz, fz1 = -1.f;
        if (mine) {
          const float r = sqrtf(__uint_as_float(best) * 1.0001f) * h.inv_s * 1.0001f + 1e-3f;  // cells (inf: everything)
          const float ry = r + 1e-4f + 1e-6f * (fabsf(uy) + (float)gy), rz = r + 1e-4f + 1e-6f * (fabsf(uz) + (float)gz);
          fy0 = fminf(fmaxf(floorf(uy - ry) - 1.f, 0.f), (float)(gy - 1));
          fy1 = fminf(fmaxf(floorf(uy + ry) + 1.f, 0.f), (float)(gy - 1));
          fz0 = fminf(fmaxf(floorf(uz - rz) - 1.f, 0.f), (float)(gz - 1));
          fz1 = fminf(fmaxf(floorf(uz + rz) + 1.f, 0.f), (float)(gz - 1));
          if (!(r < inf)) fy0 = 0.f, fy1 = (float)(gy - 1), fz0 = 0.f, fz1 = (float)(gz - 1);
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
          fy0 = fminf(fy0, __shfl_xor_sync(0xffffffffu, fy0, off));
          fy1 = fmaxf(fy1, __shfl_xor_sync(0xffffffffu, fy1, off));
          fz0 = fminf(fz0, __shfl_xor_sync(0xffffffffu, fz0, off));
          fz1 = fmaxf(fz1, __shfl_xor_sync(0xffffffffu, fz1, off));
        }
        ylo = (int)fy0, yhi = (int)fy1, zlo = (int)fz0, zhi = (int)fz1;
      }
      const int ny = yhi - ylo + 1, nrows = ny * (zhi - zlo + 1);
      for (int r0 = 0; r0 < nrows; r0 += 32) {
        // ---- lanes over rows: does any query need my row, and which cells of it?
        int a = 0, len = 0;
        unsigned need = 0;  // the queries whose bound reaches my row
        const int row = r0 + lane;
        if (row < nrows) {
          if (!h.valid) {
            a = 0, len = nt, need = sub;
          } else {
            const int yy = ylo + row % ny, zz = zlo + row / ny;
            float fx0 = (float)gx, fx1 = -1.f;
            unsigned mq = sub;
#pragma unroll 1
            while (mq) {
              const int l = __ffs(mq) - 1;
              mq &= mq - 1;
              const float bnd = sQ[l].w;
              const float4 u = sU[l];
              const float gyy = cell_gap(u.y, yy, 1e-4f + 1e-6f * (fabsf(u.y) + (float)gy)) * h.s;
              const float gzz = cell_gap(u.z, zz, 1e-4f + 1e-6f * (fabsf(u.z) + (float)gz)) * h.s;
              const float lbyz = fmaf(gyy, gyy, gzz * gzz);
              if (!(lbyz * shr > bnd)) {
                // cells c of the row with (gap_x(c) s)^2 + lbyz <= bound / shr — a superset (the square root rounded
                // up, one extra cell at either end).  In world units: the gap in CELLS of a far query against a tiny
                // cloud is astronomically large and its square overflows, the distance itself does not.
                const float xr = sqrtf(fmaxf(bnd * 1.0001f - lbyz * shr, 0.f)) * h.inv_s * 1.0001f + 1e-4f +
                                 1e-6f * (fabsf(u.x) + (float)gx) + 1e-3f;
                float f0 = fminf(fmaxf(floorf(u.x - xr) - 1.f, 0.f), (float)(gx - 1));
                float f1 = fminf(fmaxf(floorf(u.x + xr) + 1.f, 0.f), (float)(gx - 1));
                if (!(xr < inf)) f0 = 0.f, f1 = (float)(gx - 1);
                fx0 = fminf(fx0, f0), fx1 = fmaxf(fx1, f1);
                need |= 1u << l;
              }
            }
            if (fx1 >= fx0) {
              const int base = (zz * gy + yy) * gx;
              a = __ldg(start + base + (int)fx0);
              len = __ldg(start + base + (int)fx1 + 1) - a;
            }
          }
        }
        int incl = len;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        __syncwarp();
        sOff[lane] = incl - len, sPos[lane] = a;
        // ---- lanes over candidates, one tile of the batch's list at a time
        for (int t0 = 0; t0 < total; t0 += kDqTile) {
          const int navail = min(total - t0, kDqTile);
          __syncwarp();
          dq_stage<32>(C, T, sOff, sPos, t0, navail, lane);
          __syncwarp();
          // only the queries that need one of the rows this tile holds (the 32 queries of a warp may look at the
          // target from different sides)
          const bool in_tile = len > 0 && incl - len < t0 + navail && incl > t0;
          unsigned qmask;
          asm volatile("redux.sync.or.b32 %0, %1, 0xffffffff;" : "=r"(qmask) : "r"(in_tile ? need : 0u));
          evaluate(navail, qmask & sub);
        }
        __syncwarp();
        if (active) sQ[lane].w = __uint_as_float(best);  // the tightened bound prunes the following batches
        __syncwarp();
      }
      }  // classes
    }
    if (present) {
      dist[orig] = __uint_as_float(best);
      idx[orig] = bidx == 0x7fffffff ? 0 : bidx;
    }
  }
}

int chamfer_rest_launch(int b, int n, int m, const GridWs &W, const float *xyz1, const float *xyz2, float *dist1,
                        float *dist2, int *idx1, int *idx2, cudaStream_t s) {
  chamfer_rest_kernel<<<dim3(kDrCtasPerList, 2 * b), kDrThreads, 0, s>>>(b, n, m, W, xyz1, xyz2, dist1, dist2, idx1, idx2);
  count_launch();
  return launch_status();
}

}  // namespace mvp
