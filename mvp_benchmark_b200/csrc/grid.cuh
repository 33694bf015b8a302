// Uniform grid over one point cloud: header + cell addressing shared by the grid-pruned Chamfer path
// (chamfer_grid.cu) and the grid-pruned EMD bid search (emd.cu).
#pragma once
#include "common.cuh"

namespace mvp {

struct __align__(16) GridHdr {
  float lo[3];   // lower corner of the bounding box
  float inv_s;   // 1 / cell side
  float s;       // cell side (isotropic)
  int g[3];      // cells per axis; x is the fastest-varying cell coordinate
  int ncell;
  int valid;     // 0: non-finite coordinates or degenerate extent -> one cell, no pruning possible
  int pad[6];    // pad[0] = 1: every point of the cloud is the same (finite) point
};
static_assert(sizeof(GridHdr) == 64, "GridHdr layout");

// cell coordinate of u = (x - lo) * inv_s: floor + clamp in float first (u may be far outside the int range for a
// query outside the grid)
__device__ __forceinline__ int cell_coord(float u, int g) {
  return (int)fminf(fmaxf(floorf(u), 0.f), (float)(g - 1));
}

// Distance (in cells, >= 0, conservative) from coordinate u to the slab of cell c.  `slack` absorbs the rounding of
// u = fl(fl(x - lo) * inv_s): both roundings are relative 2^-24, so a point whose computed u is >= k lies at
// x - lo >= k / inv_s * (1 - 2^-23); the gap between a query at u_p and the face k is therefore at least
// (k - u_p) - 2^-23 (k + |u_p|) cells.  Callers pass slack = 1e-4 + 1e-6 (|u| + g): an 8x margin.
__device__ __forceinline__ float cell_gap(float u, int c, float slack) {
  const float g = fmaxf((float)c - u, u - (float)(c + 1));
  return fmaxf(g - slack, 0.f);
}

// Header for a cloud with bounding box [lo, hi] (`fin` = all coordinates finite): isotropic cells, at most `cap`
// of them.  Run by one thread.
__device__ inline GridHdr grid_header(const float lo[3], const float hi[3], int fin, int cap) {
  GridHdr h;
  float ex[3], emax = 0.f;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    ex[a] = hi[a] - lo[a];
    emax = fmaxf(emax, ex[a]);
    h.lo[a] = lo[a];
  }
  h.valid = (fin && emax > 0.f && emax < 1e18f) ? 1 : 0;
  h.g[0] = h.g[1] = h.g[2] = 1;
  h.s = 1.f;
  h.inv_s = 1.f;
  if (!h.valid) {
    h.lo[0] = h.lo[1] = h.lo[2] = 0.f;
  } else {
    // isotropic cell side: start from the volume heuristic (thin extents padded), grow until <= cap cells
    float vol = 1.f;
#pragma unroll
    for (int a = 0; a < 3; a++) vol *= fmaxf(ex[a], emax * 1e-3f) / emax;  // relative: no overflow
    float s = emax * cbrtf(vol / (float)cap);
    for (int it = 0; it < 400; it++) {
      float cells = 1.f;
#pragma unroll
      for (int a = 0; a < 3; a++) cells *= floorf(ex[a] / s) + 1.f;
      if (cells <= (float)cap) break;
      s *= 1.04f;
    }
    const float inv_s = 1.0f / s;
    long long cells = 1;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      // one more cell than floor(extent / s): the largest coordinate never needs the clamp by more than rounding
      const float gf = floorf(ex[a] * inv_s) + 1.f;
      h.g[a] = (int)fminf(gf, (float)cap);
      cells *= h.g[a];
    }
    if (cells > cap || !(inv_s > 0.f) || !(inv_s < 3.0e38f)) {  // pathological rounding: fall back
      h.valid = 0;
      h.g[0] = h.g[1] = h.g[2] = 1;
      h.lo[0] = h.lo[1] = h.lo[2] = 0.f;
    } else {
      h.s = s;
      h.inv_s = inv_s;
    }
  }
  h.ncell = h.g[0] * h.g[1] * h.g[2];
#pragma unroll
  for (int a = 0; a < 6; a++) h.pad[a] = 0;
  h.pad[0] = (fin && emax == 0.f) ? 1 : 0;
  return h;
}

// Workspace of the grid-pruned searches (chamfer_grid.cu: Chamfer, three_nn, mvp_knn_points), carved out of the
// caller's buffer by grid_plan(); also what the completion pass (chamfer_rest.cu) works on.
// plan words of the completion pass: two lists of work items (A: lanes over queries, B: a warp per query), each with
// its length and the number of items drawn so far; list A's items at kPlanItems, list B's rest_plan_cap() further on
constexpr int kPlanTotal = 0, kPlanTicket2 = 1, kPlanTotalB = 2, kPlanTicketB = 3, kPlanItems = 4;
inline int rest_plan_cap(int b, int n, int m) { return b * (n / 32 + m / 32 + 2); }
struct GridWs {  // carved out of the caller's workspace by grid_plan()
  GridHdr *hdr;        // [2][b]
  int *count;          // [2][b]  left-over list lengths
  int *start[2];       // [b][cap_side + 1]
  float4 *sorted[2];   // [b][n] / [b][m]
  int *list[2];        // [b][n] / [b][m]
  int cap[2];
  // completion pass (chamfer_rest.cu): bounding boxes over runs of the sorted array and the list of work items
  float4 *box[2];      // [b][2 * (nleaf + nl1 + nl2)]  (lo, hi) pairs: leaves, then level 1, then level 2
  int *plan;           // [kPlanTotal] number of work items, [kPlanTicket2] items drawn, [kPlanItems...] (list << 15) | chunk of 32 list entries
  int nleaf[2];        // leaves per cloud: ceil(points / 32)
};

}  // namespace mvp
