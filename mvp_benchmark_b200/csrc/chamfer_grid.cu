// Chamfer distance, grid-pruned exact path for sm_100a.
//
// Same outputs, bit for bit, as the brute-force kernels (chamfer.cu / chamfer_fused.cu) and therefore as the
// reference's NmDistanceKernel (utils/metrics/CD/chamfer3D/chamfer3D.cu:12-134): every candidate that is
// evaluated uses the reference's contraction  d = fma(dz,dz, fma(dx,dx, dy*dy)),  equal minima resolve to the
// lowest index (:36,:126), and a candidate is only ever SKIPPED when a conservative lower bound of its distance
// is strictly greater than a distance already found.  What changes is the amount of work: B*N*M pair
// evaluations become ~B*(N+M)*(a few dozen).
//
//   build   per (cloud, side): bounding box -> isotropic cell size with <= `cap` cells -> counting sort of the cloud
//           by cell (shared-memory histogram, block scan, scatter) into a float4 array (x, y, z, original index) +
//           a cell-start table.  x is the fastest-varying cell coordinate, so a run of cells along x is ONE
//           contiguous range of the sorted array.  One CTA (chamfer_grid_build_kernel), or a cluster of two CTAs
//           meeting through distributed shared memory (chamfer_grid_build2_kernel) for clouds of 6144-16384 points.
//   query   one thread per point, visited in the SORTED order of its own cloud (neighbouring threads look at
//           neighbouring cells of the other cloud's grid).  Cubes of cells of growing radius r around the
//           point's cell; rows of cells whose lower bound exceeds the best distance are skipped; the search
//           stops when the best distance is below the distance to the unvisited exterior.  A point that has
//           not finished after kMaxRing rings or kBudget candidates (far outside the other cloud, or inside a
//           very dense cell) is appended to a left-over list instead.
//   rest    left-over points are finished by the tiled brute-force kernel of chamfer.cu, gathering its queries
//           through the list (CTAs beyond the list length exit at once; normally all of them do).
//
// Degenerate inputs (non-finite coordinates, zero or astronomically large extent) mark the grid invalid; all
// queries against it go to the left-over list, i.e. the brute-force path and its semantics.
//
// The same build + query serve three_nn (K = 3) and the models' k-nearest-neighbour search (mvp_knn_points: an
// ordered list of up to 32 neighbours per thread, exhaustive top-k kernel for what the grid does not finish).
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "grid.cuh"

namespace mvp {

constexpr int kGridThreads = 1024;   // build CTA
constexpr int kGridMaxCells = 32768; // shared-memory histogram: 128 KB
#ifndef MVP_GRID_QTHREADS
#define MVP_GRID_QTHREADS 128
#endif
constexpr int kGridQThreads = MVP_GRID_QTHREADS;  // query CTA
#ifndef MVP_GRID_PPC
#define MVP_GRID_PPC 2               // target points per cell
#endif
#ifndef MVP_GRID_MAXRING
#define MVP_GRID_MAXRING 3
#endif
#ifndef MVP_GRID_BUDGET
#define MVP_GRID_BUDGET 512          // candidate evaluations per query before it gives up
#endif

#ifndef MVP_GRID_SCAN_UNROLL
#define MVP_GRID_SCAN_UNROLL 2       // candidate loop of the query kernel (1: 0.144, 2: 0.137, 4: 0.143 ms forward at the headline size)
#endif
constexpr int kScanUnroll = MVP_GRID_SCAN_UNROLL;

static int grid_cap(int npts) {
  int c = npts / MVP_GRID_PPC;
  c = std::max(c, 8);
  return std::min(c, kGridMaxCells);
}

static size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

static size_t grid_plan(int b, int n, int m, void *base, GridWs *w) {
  size_t off = 0;
  unsigned char *p = reinterpret_cast<unsigned char *>(base);
  auto take = [&](size_t bytes) {
    unsigned char *r = p ? p + off : nullptr;
    off += align16(bytes);
    return r;
  };
  const int cap0 = grid_cap(n), cap1 = grid_cap(m);
  GridWs t;
  t.cap[0] = cap0;
  t.cap[1] = cap1;
  t.hdr = reinterpret_cast<GridHdr *>(take(sizeof(GridHdr) * 2 * (size_t)b));
  t.count = reinterpret_cast<int *>(take(sizeof(int) * 2 * (size_t)b));
  t.start[0] = reinterpret_cast<int *>(take(sizeof(int) * (size_t)b * (cap0 + 1)));
  t.start[1] = reinterpret_cast<int *>(take(sizeof(int) * (size_t)b * (cap1 + 1)));
  t.sorted[0] = reinterpret_cast<float4 *>(take(sizeof(float4) * (size_t)b * n));
  t.sorted[1] = reinterpret_cast<float4 *>(take(sizeof(float4) * (size_t)b * m));
  t.list[0] = reinterpret_cast<int *>(take(sizeof(int) * (size_t)b * n));
  t.list[1] = reinterpret_cast<int *>(take(sizeof(int) * (size_t)b * m));
  for (int side = 0; side < 2; side++) {
    const int nleaf = ((side ? m : n) + 31) / 32, nl1 = (nleaf + 31) / 32, nl2 = (nl1 + 31) / 32;
    t.nleaf[side] = nleaf;
    t.box[side] = reinterpret_cast<float4 *>(take(sizeof(float4) * 2 * (size_t)b * (nleaf + nl1 + nl2)));
  }
  t.plan = reinterpret_cast<int *>(take(sizeof(int) * (kPlanItems + 2 * (size_t)rest_plan_cap(b, n, m))));
  if (w) *w = t;
  return off;
}

// ------------------------------------------------------------------------------------------------ build
// Clouds of up to kGridRankPts points keep (cell, rank inside the cell) of every point in shared memory between the
// histogram and the scatter pass: the rank is what the histogram's atomicAdd returned, so the scatter needs no second
// round of shared-memory atomics (2 cycles per lane each — they, not the loads, bound this kernel).
constexpr int kGridRankPts = 16384;
__host__ __device__ inline size_t grid_hist_words(const int cap[2]) {  // histogram, padded for 128-bit access
  return ((size_t)(cap[0] > cap[1] ? cap[0] : cap[1]) + 7) & ~(size_t)7;
}
static size_t grid_build_smem(const int cap[2], int n, int m) {
  const int np = std::max(n, m);
  return sizeof(int) * (grid_hist_words(cap) + (np <= kGridRankPts ? (size_t)np : 0));
}

__global__ void __launch_bounds__(kGridThreads, 1)
chamfer_grid_build_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                          GridWs W) {
  extern __shared__ __align__(16) int hist[];  // cap ints
  __shared__ float s_red[6][32];
  __shared__ int s_fin[32];
  __shared__ int s_warp[32];
  __shared__ GridHdr s_hdr;
  pdl_trigger();
  const int cloud = blockIdx.x, side = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = side ? m : n;
  const int cap = side ? W.cap[1] : W.cap[0];
  const float *P = (side ? xyz2 : xyz1) + (size_t)cloud * np * 3;
  const float inf = __int_as_float(0x7f800000);

  // ---- pass 1: bounding box and finiteness
  float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  int fin = 1;
  // every pass walks the cloud in batches of kBatch points per thread with all loads of a batch issued before their
  // first use: the passes are latency-bound (one CTA per SM, 16 points per thread at n = 16384)
  constexpr int kBatch = 4;
  for (int i0 = tid; i0 < np; i0 += kBatch * kGridThreads) {
    float v[kBatch][3];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int i = min(i0 + u * kGridThreads, np - 1);  // clamped: a repeated point changes neither box nor flag
#pragma unroll
      for (int a = 0; a < 3; a++) v[u][a] = __ldg(P + (size_t)i * 3 + a);
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        lo[a] = fminf(lo[a], v[u][a]);
        hi[a] = fmaxf(hi[a], v[u][a]);
        fin &= (fabsf(v[u][a]) <= 3.0e38f) ? 1 : 0;  // false for NaN and +-inf
      }
    }
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
    for (int a = 0; a < 3; a++) {
      s_red[a][warp] = lo[a];
      s_red[3 + a][warp] = hi[a];
    }
    s_fin[warp] = fin;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kGridThreads / 32; w++) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        lo[a] = fminf(lo[a], s_red[a][w]);
        hi[a] = fmaxf(hi[a], s_red[3 + a][w]);
      }
      fin &= s_fin[w];
    }
    const GridHdr h = grid_header(lo, hi, fin, cap);
    s_hdr = h;
    W.hdr[side * b + cloud] = h;
    W.count[side * b + cloud] = 0;
    if (cloud == 0 && side == 0) W.plan[kPlanTotal] = 0, W.plan[kPlanTicket2] = 0, W.plan[kPlanTotalB] = 0, W.plan[kPlanTicketB] = 0;
  }
  __syncthreads();
  const GridHdr h = s_hdr;
  const int ncell = h.ncell;
  for (int c = tid; c < ncell; c += kGridThreads) hist[c] = 0;
  __syncthreads();

  auto cell_of = [&](float x, float y, float z) {
    if (!h.valid) return 0;
    const int cx = cell_coord((x - h.lo[0]) * h.inv_s, h.g[0]);
    const int cy = cell_coord((y - h.lo[1]) * h.inv_s, h.g[1]);
    const int cz = cell_coord((z - h.lo[2]) * h.inv_s, h.g[2]);
    return (cz * h.g[1] + cy) * h.g[0] + cx;
  };

  // ---- pass 2: histogram (and, for clouds that fit, every point's cell and arrival rank)
  const bool ranked = max(n, m) <= kGridRankPts;
  unsigned *code = reinterpret_cast<unsigned *>(hist + grid_hist_words(W.cap));  // [np] when ranked
  for (int i0 = tid; i0 < np; i0 += kBatch * kGridThreads) {
    float v[kBatch][3];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int i = min(i0 + u * kGridThreads, np - 1);
#pragma unroll
      for (int a = 0; a < 3; a++) v[u][a] = __ldg(P + (size_t)i * 3 + a);
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int i = i0 + u * kGridThreads;
      if (i < np) {
        const int c = cell_of(v[u][0], v[u][1], v[u][2]);
        const int r = atomicAdd(&hist[c], 1);
        if (ranked) code[i] = ((unsigned)c << 15) | (unsigned)r;  // c < 2^15 cells, r < 2^14 points
      }
    }
  }
  __syncthreads();

  // ---- exclusive scan over the cells (contiguous chunk per thread, warp scan, scan of warp totals)
  const int per = (ncell + kGridThreads - 1) / kGridThreads;
  const int c0 = min(tid * per, ncell), c1 = min(c0 + per, ncell);
  int sum = 0;
  for (int c = c0; c < c1; c++) sum += hist[c];
  int incl = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = s_warp[lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, off);
      if (lane >= off) v += u;
    }
    s_warp[lane] = v;
  }
  __syncthreads();
  int run = incl - sum + (warp ? s_warp[warp - 1] : 0);
  int *start = (side ? W.start[1] : W.start[0]) + (size_t)cloud * (cap + 1);
  for (int c = c0; c < c1; c++) {
    const int cnt = hist[c];
    hist[c] = run;  // becomes the fill cursor of the scatter pass
    start[c] = run;
    run += cnt;
  }
  if (tid == 0) start[ncell] = np;
  __syncthreads();
  // ---- pass 3: scatter (order inside a cell is arbitrary; the query's tie rule is explicit)
  float4 *S = (side ? W.sorted[1] : W.sorted[0]) + (size_t)cloud * np;
  for (int i0 = tid; i0 < np; i0 += kBatch * kGridThreads) {
    float v[kBatch][3];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int i = min(i0 + u * kGridThreads, np - 1);
#pragma unroll
      for (int a = 0; a < 3; a++) v[u][a] = __ldg(P + (size_t)i * 3 + a);
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int i = i0 + u * kGridThreads;
      if (i < np) {
        int pos;
        if (ranked) {
          const unsigned cr = code[i];
          pos = hist[cr >> 15] + (int)(cr & 0x7fffu);  // hist holds the cell starts now
        } else {
          pos = atomicAdd(&hist[cell_of(v[u][0], v[u][1], v[u][2])], 1);
        }
        S[pos] = make_float4(v[u][0], v[u][1], v[u][2], __int_as_float(i));
      }
    }
  }
}

// ---- the same build for clouds of up to kGridRankPts points, split over a cluster of TWO CTAs ------------------------
// One CTA per (cloud, side) leaves 84 of the 148 SMs idle at B = 32 and every thread walks 16 points through three
// dependent passes.  Here each CTA of a pair owns half of the points: partial bounding boxes and partial histograms
// meet through distributed shared memory (two cluster barriers), both CTAs derive the same header and the same
// cell starts, and CTA 1's cursor of a cell starts behind CTA 0's points of that cell.
#ifdef MVP_GRID_BUILD_TIMING  // debugging aid: per-phase clock64 stamps of thread 0, printed by cloud 0 / side 0
#define MVP_B2_STAMP(k) b2_t[k] = clock64()
#else
#define MVP_B2_STAMP(k)
#endif
constexpr int kBuild2MinPts = 6144;  // measured (tools/grid_sizes.py): at 3072 points one CTA per side is 7-9 us faster, at 16384 10-14 us slower
constexpr int kBuild2Per = 8;  // cells per thread: clouds whose grids have up to 8192 cells (16384 points at 2 per cell)

__global__ void __cluster_dims__(1, 1, 2) __launch_bounds__(kGridThreads, 1)
chamfer_grid_build2_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                           GridWs W) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) int hist[];  // cap ints, then the codes of this CTA's points
  __shared__ float s_red[6][32];
  __shared__ int s_fin[32];
  __shared__ int s_warp[32];
  __shared__ float s_part[8];    // this CTA's box (lo xyz, hi xyz) and finiteness flag, read by the peer
  __shared__ GridHdr s_hdr;
  pdl_trigger();
  const int cloud = blockIdx.x, side = blockIdx.y;
  const int rank = (int)cluster.block_rank(), peer = rank ^ 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = side ? m : n;
  const int cap = side ? W.cap[1] : W.cap[0];
  const float *P = (side ? xyz2 : xyz1) + (size_t)cloud * np * 3;
  const int half0 = (np + 1) / 2;
  const int first = rank ? half0 : 0, mine = rank ? np - half0 : half0;  // this CTA's points: [first, first + mine)
  const float inf = __int_as_float(0x7f800000);
  constexpr int kBatch = 4;
#ifdef MVP_GRID_BUILD_TIMING
  long long b2_t[10];
#endif
  MVP_B2_STAMP(0);

  // ---- pass 1: bounding box and finiteness of this half
  float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  int fin = 1;
  for (int j0 = tid; j0 < mine; j0 += kBatch * kGridThreads) {
    float v[kBatch][3];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int i = first + min(j0 + u * kGridThreads, mine - 1);
#pragma unroll
      for (int a = 0; a < 3; a++) v[u][a] = __ldg(P + (size_t)i * 3 + a);
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        lo[a] = fminf(lo[a], v[u][a]);
        hi[a] = fmaxf(hi[a], v[u][a]);
        fin &= (fabsf(v[u][a]) <= 3.0e38f) ? 1 : 0;
      }
    }
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
    for (int a = 0; a < 3; a++) {
      s_red[a][warp] = lo[a];
      s_red[3 + a][warp] = hi[a];
    }
    s_fin[warp] = fin;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kGridThreads / 32; w++) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        lo[a] = fminf(lo[a], s_red[a][w]);
        hi[a] = fmaxf(hi[a], s_red[3 + a][w]);
      }
      fin &= s_fin[w];
    }
#pragma unroll
    for (int a = 0; a < 3; a++) s_part[a] = lo[a], s_part[3 + a] = hi[a];
    s_part[6] = __int_as_float(fin);
  }
  MVP_B2_STAMP(1);
  cluster.sync();
  MVP_B2_STAMP(2);
  if (tid == 0) {
    const float *q = cluster.map_shared_rank(s_part, peer);
#pragma unroll
    for (int a = 0; a < 3; a++) {  // min / max commute: both CTAs arrive at the same box, hence the same header
      lo[a] = fminf(lo[a], q[a]);
      hi[a] = fmaxf(hi[a], q[3 + a]);
    }
    fin &= __float_as_int(q[6]);
    const GridHdr h = grid_header(lo, hi, fin, cap);
    s_hdr = h;
    if (rank == 0) {
      W.hdr[side * b + cloud] = h;
      W.count[side * b + cloud] = 0;
      if (cloud == 0 && side == 0) W.plan[kPlanTotal] = 0, W.plan[kPlanTicket2] = 0, W.plan[kPlanTotalB] = 0, W.plan[kPlanTicketB] = 0;
    }
  }
  __syncthreads();
  const GridHdr h = s_hdr;
  const int ncell = h.ncell;
  MVP_B2_STAMP(3);
  for (int c = tid; c < ncell; c += kGridThreads) hist[c] = 0;
  __syncthreads();
  MVP_B2_STAMP(4);

  auto cell_of = [&](float x, float y, float z) {
    if (!h.valid) return 0;
    const int cx = cell_coord((x - h.lo[0]) * h.inv_s, h.g[0]);
    const int cy = cell_coord((y - h.lo[1]) * h.inv_s, h.g[1]);
    const int cz = cell_coord((z - h.lo[2]) * h.inv_s, h.g[2]);
    return (cz * h.g[1] + cy) * h.g[0] + cx;
  };

  // ---- pass 2: histogram of this half; every point keeps (cell, arrival rank) for the scatter
  unsigned *code = reinterpret_cast<unsigned *>(hist + grid_hist_words(W.cap));  // [mine]
  for (int j0 = tid; j0 < mine; j0 += kBatch * kGridThreads) {
    float v[kBatch][3];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int i = first + min(j0 + u * kGridThreads, mine - 1);
#pragma unroll
      for (int a = 0; a < 3; a++) v[u][a] = __ldg(P + (size_t)i * 3 + a);
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int j = j0 + u * kGridThreads;
      if (j < mine) {
        const int c = cell_of(v[u][0], v[u][1], v[u][2]);
        code[j] = ((unsigned)c << 15) | (unsigned)atomicAdd(&hist[c], 1);
      }
    }
  }
  MVP_B2_STAMP(5);
  cluster.sync();  // both partial histograms are complete
  MVP_B2_STAMP(6);

  // ---- exclusive scan over the cells of (own + peer) counts; the counts sit in registers across the barrier after
  // which this CTA overwrites its histogram with its cursors
  // (each thread owns kBuild2Per = 8 consecutive cells: two 128-bit loads from each histogram — scalar loads at a
  // stride of 8 words are 8-way bank conflicts here and eight remote round trips there; measured 19-24k of the
  // kernel's 47k cycles before)
  static_assert(kBuild2Per == 8, "two int4 per thread");
  const int4 *own4 = reinterpret_cast<const int4 *>(hist);
  const int4 *oth4 = reinterpret_cast<const int4 *>(cluster.map_shared_rank(hist, peer));
  const int c0 = tid * kBuild2Per;
  int own[kBuild2Per], oth[kBuild2Per], sum = 0;
  {
    int4 a0 = make_int4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
    if (c0 < ncell) {  // words past ncell (inside the padded allocation) are masked below
      a0 = own4[2 * tid], a1 = own4[2 * tid + 1];
      b0 = oth4[2 * tid], b1 = oth4[2 * tid + 1];
    }
    const int ta[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const int tb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int e = 0; e < kBuild2Per; e++) {
      own[e] = c0 + e < ncell ? ta[e] : 0;
      oth[e] = c0 + e < ncell ? tb[e] : 0;
      sum += own[e] + oth[e];
    }
  }
  int incl = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  cluster.sync();  // (also a CTA barrier) the peer has read this CTA's histogram: it may be overwritten now
  if (warp == 0) {
    int v = s_warp[lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, off);
      if (lane >= off) v += u;
    }
    s_warp[lane] = v;
  }
  __syncthreads();
  int run = incl - sum + (warp ? s_warp[warp - 1] : 0);
  if (c0 < ncell) {
    int cur[kBuild2Per];
#pragma unroll
    for (int e = 0; e < kBuild2Per; e++) {
      cur[e] = run + (rank ? oth[e] : 0);  // CTA 1 fills a cell behind CTA 0's points
      run += own[e] + oth[e];
    }
    int4 *h4 = reinterpret_cast<int4 *>(hist);
    h4[2 * tid] = make_int4(cur[0], cur[1], cur[2], cur[3]);
    h4[2 * tid + 1] = make_int4(cur[4], cur[5], cur[6], cur[7]);
  }
  __syncthreads();
  int *start = (side ? W.start[1] : W.start[0]) + (size_t)cloud * (cap + 1);
  if (rank == 0)  // CTA 0's cursors are the cell starts: coalesced copy
    for (int c = tid; c < ncell; c += kGridThreads) start[c] = hist[c];
  if (rank == 0 && tid == 0) start[ncell] = np;
  __syncthreads();
  MVP_B2_STAMP(7);

  MVP_B2_STAMP(8);
  // ---- pass 3: scatter this half (order inside a cell is arbitrary; the query's tie rule is explicit)
  float4 *S = (side ? W.sorted[1] : W.sorted[0]) + (size_t)cloud * np;
  for (int j0 = tid; j0 < mine; j0 += kBatch * kGridThreads) {
    float v[kBatch][3];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int i = first + min(j0 + u * kGridThreads, mine - 1);
#pragma unroll
      for (int a = 0; a < 3; a++) v[u][a] = __ldg(P + (size_t)i * 3 + a);
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int j = j0 + u * kGridThreads;
      if (j < mine) {
        const unsigned cr = code[j];
        S[hist[cr >> 15] + (int)(cr & 0x7fffu)] = make_float4(v[u][0], v[u][1], v[u][2], __int_as_float(first + j));
      }
    }
  }
  MVP_B2_STAMP(9);
#ifdef MVP_GRID_BUILD_TIMING
  __syncthreads();
  if (tid == 0 && cloud == 0 && side == 0)
    printf("build2 rank %d: pass1 %lld sync %lld header %lld zero %lld pass2 %lld sync %lld scan %lld keys %lld pass3 %lld total %lld\n", rank,
           b2_t[1] - b2_t[0], b2_t[2] - b2_t[1], b2_t[3] - b2_t[2], b2_t[4] - b2_t[3], b2_t[5] - b2_t[4], b2_t[6] - b2_t[5],
           b2_t[7] - b2_t[6], b2_t[8] - b2_t[7], b2_t[9] - b2_t[8], b2_t[9] - b2_t[0]);
#endif
}

// Launches the build for both sides: the two-CTA cluster kernel when both clouds fit it, else one CTA per side.
static int grid_build_launch(int b, int n, int m, const float *xyz1, const float *xyz2, const GridWs &W,
                             cudaStream_t s) {
  const size_t smem = grid_build_smem(W.cap, n, m);
  {
    static size_t granted[2][kMaxDevices];
    int rc0 = grant_dyn_smem(chamfer_grid_build_kernel, sizeof(int) * (kGridMaxCells + 8 + kGridRankPts), granted[0], 0);
    if (!rc0)
      rc0 = grant_dyn_smem(chamfer_grid_build2_kernel, sizeof(int) * (kBuild2Per * kGridThreads + 8 + kGridRankPts),
                           granted[1], 0);
    if (rc0) return rc0;
  }
  static const int min_pts = [] {  // tuning aid: clouds below MVP_GRID_BUILD2_MIN points use one CTA per side
    const char *e = getenv("MVP_GRID_BUILD2_MIN");
    return e ? atoi(e) : kBuild2MinPts;
  }();
  if (std::max(n, m) >= min_pts && std::max(n, m) <= kGridRankPts &&
      std::max(W.cap[0], W.cap[1]) <= kBuild2Per * kGridThreads)
    chamfer_grid_build2_kernel<<<dim3(b, 2, 2), kGridThreads, smem, s>>>(b, n, m, xyz1, xyz2, W);
  else
    chamfer_grid_build_kernel<<<dim3(b, 2), kGridThreads, smem, s>>>(b, n, m, xyz1, xyz2, W);
  return MVP_OK;
}

// ------------------------------------------------------------------------------------------------ query
// K = 1: Chamfer, both directions (points of xyz1 against xyz2's grid, then the other way round).
// K = 3: three_nn — the n points of side 0 against side 1's grid only; the three nearest in ascending (distance,
//        index) order, which is what the reference's strict-`<` scan in index order keeps (three_nn_cuda.cu:43-58);
//        outputs (b, n, 3).  Pruning and termination then test the THIRD best distance.
// kRT (top-k nearest neighbours, mvp_knn_points): K is the register capacity, the runtime `kk` <= K the number of
//        neighbours wanted.  The first K - kk slots of the ordered list are phantoms at distance -inf, so that the
//        kk-th real neighbour is always slot K - 1 — every array index stays a compile-time constant (a runtime
//        `bd[kk - 1]`, even written as a chain of selects, sends the list to local memory).  Outputs are (b, n, kk).
//        Larger candidate budget and one more ring than the K <= 3 searches.
template <int K>
__device__ __forceinline__ void topk_insert(float (&bd)[K], int (&bk)[K], float d, int qi) {
  bd[K - 1] = d;
  bk[K - 1] = qi;
#pragma unroll
  for (int k = K - 1; k > 0; k--) {
    if (bd[k] < bd[k - 1] || (bd[k] == bd[k - 1] && bk[k] < bk[k - 1])) {
      const float td_ = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td_;
      const int tk_ = bk[k]; bk[k] = bk[k - 1]; bk[k - 1] = tk_;
    }
  }
}
#ifndef MVP_GRID_QMINB
#define MVP_GRID_QMINB 8             // resident CTAs per SM the Chamfer query kernel is compiled for.  The per-lane kernels of round 1 wanted 12 (40 registers, a few spills, more warps in flight: 0.137 -> 0.133 ms); the pooled kernel keeps more state live and measures best at 8 (64 registers, no spills): forward 0.1055 -> 0.1029 ms at the headline size, 0.0416 -> 0.0385 ms at 64 x 2048 x 2048
#endif
#ifndef MVP_GRID_QMINB3
#define MVP_GRID_QMINB3 12           // the same for three_nn (K = 3): 68 -> 63 us at 64 x 3072 from 1536; 0 = no occupancy target
#endif
#ifndef MVP_GRID_QMINBK
#define MVP_GRID_QMINBK 6            // ... and for the k-nearest-neighbour lists of 12 and 16 entries (356 -> 319 us at 64 x 3072, k = 16; 8: 339)
#endif
// kDyn (Chamfer only): the nine rows of the 3x3x3 cube are not walked in lockstep.  Unrolled, every row's set-up and
// candidate loop is executed by the whole warp as soon as ONE lane needs the row: ncu (round 2) shows the four face
// rows running with 7-14 of 32 lanes and the four diagonal rows with 3-6, 70 % of the kernel's instructions.  Here
// every lane advances to ITS next row that its current bound does not prune (same fixed order, same progressive
// pruning, hence the same candidates and results), so a warp runs for the largest NUMBER of rows a lane needs —
// about five or six — instead of nine.
// kPool (with kDyn): the (lane, row) pairs a warp has left after the centre row are POOLED and dealt out 32 at a time —
// see the ring-1 block.
struct QueryPool {  // per warp
  float4 s0[32];                 // self x y z, best after the centre row
  float4 s1[32];                 // gys[0] gys[2] gzs[0] gzs[2]
  float4 s2[32];                 // gxs[0] gxs[2], cx cy (as int bits)
  int cz[32];
  unsigned long long key[32];    // (bits(best) << 32) | index: the queries' running results, merged with atomicMin
  int over[32];                  // a row of this query exceeded the candidate budget
  unsigned char item[256];       // ((row - 1) << 5) | lane, in row-major order
};
template <int K, bool kRT = false, bool kDyn = false, bool kPool = false>
__global__ void __launch_bounds__(kGridQThreads, (kRT ? (K == 16 || K == 12 ? MVP_GRID_QMINBK : 0) : K == 1 ? MVP_GRID_QMINB : MVP_GRID_QMINB3))  // (an explicit 1 lets ptxas take 100+ registers)
chamfer_grid_query_kernel(int b, int n, int m, GridWs W, float *__restrict__ dist1, float *__restrict__ dist2,
                          int *__restrict__ idx1, int *__restrict__ idx2, int kk) {
  __shared__ QueryPool s_pool[kPool ? kGridQThreads / 32 : 1];
  constexpr int kBudget = kRT ? MVP_GRID_BUDGET + 64 * K : MVP_GRID_BUDGET;
  constexpr int kMaxRing = kRT ? MVP_GRID_MAXRING + 1 : MVP_GRID_MAXRING;
  pdl_wait();     // (the grid build's output; a no-op unless launched with launch_pdl)
  pdl_trigger();
  const int ko = kRT ? kk : K;  // neighbours per query in the outputs
  // (b * (n + m) < 2^31 is a precondition of every grid path: 32-bit index arithmetic, one unsigned division)
  const unsigned total1 = (unsigned)b * (unsigned)n, total = K == 1 ? total1 + (unsigned)b * (unsigned)m : total1;
  const unsigned t = blockIdx.x * (unsigned)kGridQThreads + threadIdx.x;
  if (t < total) {
  const int dir = t >= total1 ? 1 : 0;       // 0: points of xyz1 against xyz2's grid; 1: the other way round
  const unsigned pi = dir ? t - total1 : t;
  const int nq = dir ? m : n, nt = dir ? n : m;
  const int cloud = (int)(pi / (unsigned)nq);
  const float4 self = __ldg((dir ? W.sorted[1] : W.sorted[0]) + pi);
  const int orig = __float_as_int(self.w);
  const int ts = 1 - dir;  // target side
  const GridHdr *hp = W.hdr + ts * b + cloud;
  const int4 h0 = __ldg(reinterpret_cast<const int4 *>(hp));      // lo.xyz, inv_s
  const int4 h1 = __ldg(reinterpret_cast<const int4 *>(hp) + 1);  // s, g.xyz
  const int4 h2 = __ldg(reinterpret_cast<const int4 *>(hp) + 2);  // ncell, valid
  const float inv_s = __int_as_float(h0.w), s = __int_as_float(h1.x);
  const int gx = h1.y, gy = h1.z, gz = h1.w;
  const int *start = (ts ? W.start[1] : W.start[0]) + (size_t)cloud * ((ts ? W.cap[1] : W.cap[0]) + 1);
  const float4 *T = (ts ? W.sorted[1] : W.sorted[0]) + (size_t)cloud * nt;
  float *dist = (dir ? dist2 : dist1) + (size_t)cloud * nq * ko;
  int *idx = (dir ? idx2 : idx1) + (size_t)cloud * nq * ko;

  const float inf = __int_as_float(0x7f800000);
  // the K best so far, ascending in (distance, index); `best` is the one pruning tests (the K-th)
  float bd[K];
  int bk[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const bool phantom = kRT && k < K - kk;
    bd[k] = phantom ? -inf : inf;
    bk[k] = phantom ? -1 : 0x7fffffff;
  }
  float &best = bd[K - 1];
  int &bestk = bk[K - 1];
  bool done = false;
  if (h2.y) {
    const float ux = (self.x - __int_as_float(h0.x)) * inv_s;
    const float uy = (self.y - __int_as_float(h0.y)) * inv_s;
    const float uz = (self.z - __int_as_float(h0.z)) * inv_s;
    const int cx = cell_coord(ux, gx), cy = cell_coord(uy, gy), cz = cell_coord(uz, gz);
    // rounding slack of the cell coordinates, in cells (derivation in DESIGN.md §4.1): 2^-23 (|u| + g), 8x margin
    const float slx = 1e-4f + 1e-6f * (fabsf(ux) + (float)gx);
    const float sly = 1e-4f + 1e-6f * (fabsf(uy) + (float)gy);
    const float slz = 1e-4f + 1e-6f * (fabsf(uz) + (float)gz);
    const float s2 = s * s * (1.f - 1e-5f);  // (cells -> squared distance), rounded DOWN generously
    // A query more than kMaxRing + 1 cells outside the grid cannot finish: its best distance stays above the
    // lateral extent of the largest cube (ext below).  Also catches overflowing / non-finite cell coordinates.
    const float out = fmaxf(fmaxf(fmaxf(-ux, ux - (float)gx), fmaxf(-uy, uy - (float)gy)), fmaxf(-uz, uz - (float)gz));
    const bool finite = fabsf(ux) + fabsf(uy) + fabsf(uz) < 3.0e38f;  // false for NaN / inf
#ifdef MVP_GRID_NOBAIL
    int budget = kBudget;
#else
    int budget = (finite && out <= (float)(kMaxRing + 1)) ? kBudget : -1;
#endif

    // distance (in cells, >= 0, conservative) from coordinate u to the slab of cell c
    auto gap = [](float u, int c, float slack) {
      const float g = fmaxf((float)c - u, u - (float)(c + 1));
      return fmaxf(g - slack, 0.f);
    };
    auto scan = [&](int a, int e) {
      if (e - a > budget) {  // too dense here: give up, the brute-force pass finishes this point
        budget = -1;
        return;
      }
      budget -= e - a;
#pragma unroll(K == 1 ? kScanUnroll : 1)  // longer ordered lists: the duplicated insertion code costs more than it saves
      for (unsigned i = (unsigned)a; i < (unsigned)e; i++) {  // (unsigned: the address is one IMAD.WIDE.U32)
        const float4 q = __ldg(T + i);
        const float d = sqdist(q.x - self.x, q.y - self.y, q.z - self.z);
        if (d <= best) {  // rare after the first few candidates
          const int qi = __float_as_int(q.w);
          if (d < best || qi < bestk) {  // (d, qi) precedes the K-th (kk-th) best: insert it in order
            topk_insert<K>(bd, bk, d, qi);
          }
        }
      }
    };
    // cells [x0, x1] of row (yy, zz), clipped at both ends by the per-cell bound (rings >= 2 only)
    auto row = [&](int x0, int x1, int yy, int zz, float lbyz) {
      x0 = max(x0, 0);
      x1 = min(x1, gx - 1);
      while (x0 <= x1) {
        const float g = gap(ux, x0, slx);
        if (fmaf(g, g, lbyz) * s2 > best) x0++; else break;
      }
      while (x1 >= x0) {
        const float g = gap(ux, x1, slx);
        if (fmaf(g, g, lbyz) * s2 > best) x1--; else break;
      }
      if (x0 > x1) return;
      const int base = (zz * gy + yy) * gx;
      scan(__ldg(start + base + x0), __ldg(start + base + x1 + 1));
    };

    for (int r = 1; r <= kMaxRing && !done && budget >= 0; r++) {
      if (r == 1) {
        // ---- the 3x3x3 cube, unrolled: per-axis squared gaps of the two outer slabs once (inf = outside the
        // grid), the centre row first so that a good `best` prunes the other eight
        float gxs[3], gys[3], gzs[3];
        gxs[1] = gys[1] = gzs[1] = 0.f;
        { const float g = gap(ux, cx - 1, slx); gxs[0] = cx > 0 ? g * g : inf; }
        { const float g = gap(ux, cx + 1, slx); gxs[2] = cx + 1 < gx ? g * g : inf; }
        { const float g = gap(uy, cy - 1, sly); gys[0] = cy > 0 ? g * g : inf; }
        { const float g = gap(uy, cy + 1, sly); gys[2] = cy + 1 < gy ? g * g : inf; }
        { const float g = gap(uz, cz - 1, slz); gzs[0] = cz > 0 ? g * g : inf; }
        { const float g = gap(uz, cz + 1, slz); gzs[2] = cz + 1 < gz ? g * g : inf; }
        const int xlo = cx > 0 ? cx - 1 : cx, xhi = cx + 1 < gx ? cx + 1 : cx;
        if constexpr (kDyn) {
          // rows in the order centre, four faces, four diagonals: offsets + 1, two bits per row
          constexpr unsigned kOY = 1u | 0u << 2 | 2u << 4 | 1u << 6 | 1u << 8 | 0u << 10 | 2u << 12 | 0u << 14 | 2u << 16;
          constexpr unsigned kOZ = 1u | 1u << 2 | 1u << 4 | 0u << 6 | 2u << 8 | 0u << 10 | 0u << 12 | 2u << 14 | 2u << 16;
          auto sel3 = [](const float (&a)[3], int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); };
          // The loop is kept WARP-UNIFORM (one ballot per round): lanes that have run out of rows idle inside it, so
          // that the warp reconverges at the top of every round — left to themselves, diverged lanes never do
          // (measured: 15 rounds per warp and 8.7 lanes per instruction with per-lane `break`s).
          const unsigned wmask = __activemask();
          // centre row: every lane, in lockstep
          {
            const int base = (cz * gy + cy) * gx;
            scan(__ldg(start + base + xlo), __ldg(start + base + xhi + 1));
          }
          // the other eight rows: those the centre row's bound leaves, as a bit mask (all lanes, no divergence) ...
          unsigned todo = 0;
#pragma unroll
          for (int t = 1; t < 9; t++) {
            constexpr int oyt[9] = {0, -1, 1, 0, 0, -1, 1, -1, 1};
            constexpr int ozt[9] = {0, 0, 0, -1, 1, -1, -1, 1, 1};
            const float lb = gys[oyt[t] + 1] + gzs[ozt[t] + 1];
            if (!(lb * s2 > best) && lb < inf) todo |= 1u << t;  // (lb = inf: outside the grid)
          }
          if (budget < 0) todo = 0;
          // Pooled: a round of the loop below runs with ~13 of 32 lanes (a lane needs 2.7 rows on average, the
          // unluckiest of a warp six).  When the whole warp looks at ONE target grid, the (lane, row) pairs are
          // instead written to a list and dealt out 32 at a time: the lane that gets a pair fetches that query's state
          // from shared memory, clips and scans the row, and merges its result into the query's 64-bit
          // (distance, index) key with a shared-memory atomicMin — lexicographic, hence the same winner as any order
          // of evaluation.  Rows are tested against the centre row's bound only (no progressive tightening between
          // the rows of a query): a few more candidates, evaluated by full warps.
          bool pooled = false;
          if constexpr (kPool) {
            const int lane = threadIdx.x & 31;
            const int cloud0 = __shfl_sync(wmask, cloud, 0), dir0 = __shfl_sync(wmask, dir, 0);
            // one target grid for the whole warp, and a grid with rows on both sides of most cells (a planar cloud's queries
            // have two neighbour rows, not eight: not worth the hand-over)
            pooled = wmask == 0xffffffffu && gy >= 3 && gz >= 3 && __all_sync(wmask, cloud == cloud0 && dir == dir0);
            if (pooled) {
              QueryPool &P = s_pool[threadIdx.x >> 5];
              P.s0[lane] = make_float4(self.x, self.y, self.z, best);
              P.s1[lane] = make_float4(gys[0], gys[2], gzs[0], gzs[2]);
              P.s2[lane] = make_float4(gxs[0], gxs[2], __int_as_float(cx), __int_as_float(cy));
              P.cz[lane] = cz;
              P.key[lane] = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)bestk;
              P.over[lane] = 0;
              int total_items = 0;
#pragma unroll
              for (int t = 1; t < 9; t++) {
                const bool mine = (todo >> t) & 1u;
                const unsigned mt = __ballot_sync(0xffffffffu, mine);
                if (mine) P.item[total_items + __popc(mt & ((1u << lane) - 1u))] = (unsigned char)(((t - 1) << 5) | lane);
                total_items += __popc(mt);
              }
              __syncwarp();
              for (int k0 = 0; k0 < total_items; k0 += 32) {
                const int k = k0 + lane;
                if (k < total_items) {
                  const int it = P.item[k], q = it & 31, t = (it >> 5) + 1;
                  const int oy1 = (int)(kOY >> (2 * t)) & 3, oz1 = (int)(kOZ >> (2 * t)) & 3;
                  const float4 q0 = P.s0[q], q1 = P.s1[q], q2 = P.s2[q];
                  const float lby = oy1 == 1 ? 0.f : (oy1 == 0 ? q1.x : q1.y);
                  const float lbz = oz1 == 1 ? 0.f : (oz1 == 0 ? q1.z : q1.w);
                  const float lbyz = lby + lbz, bq = q0.w;
                  const int qcx = __float_as_int(q2.z), yy = __float_as_int(q2.w) + oy1 - 1, zz = P.cz[q] + oz1 - 1;
                  const int qxlo = qcx > 0 ? qcx - 1 : qcx, qxhi = qcx + 1 < gx ? qcx + 1 : qcx;
                  const int x0 = (q2.x + lbyz) * s2 > bq ? qcx : qxlo;
                  const int x1 = (q2.y + lbyz) * s2 > bq ? qcx : qxhi;
                  const int base = (zz * gy + yy) * gx;
                  const int a = __ldg(start + base + x0), e = __ldg(start + base + x1 + 1);
                  if (e - a > kBudget) {
                    P.over[q] = 1;
                  } else {
                    const float sx = q0.x, sy = q0.y, sz = q0.z;
                    float bb = bq;
                    int bi = 0x7fffffff;
#pragma unroll 2
                    for (unsigned i = (unsigned)a; i < (unsigned)e; i++) {
                      const float4 c4 = __ldg(T + i);
                      const float d = sqdist(c4.x - sx, c4.y - sy, c4.z - sz);
                      const int ci = __float_as_int(c4.w);
                      if (d < bb || (d == bb && ci < bi)) bb = d, bi = ci;
                    }
                    if (bi != 0x7fffffff)
                      atomicMin(&P.key[q], ((unsigned long long)__float_as_uint(bb) << 32) | (unsigned)bi);
                  }
                }
              }
              __syncwarp();
              const unsigned long long kq = P.key[lane];
              best = __uint_as_float((unsigned)(kq >> 32));
              bestk = (int)(unsigned)kq;
              if (P.over[lane]) budget = -1;
              todo = 0;
              __syncwarp();
            }
          }
          // ... then every lane takes ITS next such row per round and re-tests it against its current bound
          while (__any_sync(wmask, todo != 0)) {
            if (todo) {
              const int t = __ffs(todo) - 1;
              todo &= todo - 1;
              const int oy1 = (int)(kOY >> (2 * t)) & 3, oz1 = (int)(kOZ >> (2 * t)) & 3;
              const float lbyz = sel3(gys, oy1) + sel3(gzs, oz1);
              if (!(lbyz * s2 > best)) {
                const int yy = cy + oy1 - 1, zz = cz + oz1 - 1;
                const int x0 = (gxs[0] + lbyz) * s2 > best ? cx : xlo;
                const int x1 = (gxs[2] + lbyz) * s2 > best ? cx : xhi;
                const int base = (zz * gy + yy) * gx;
                scan(__ldg(start + base + x0), __ldg(start + base + x1 + 1));
                if (budget < 0) todo = 0;
              }
            }
          }
        } else
#pragma unroll
        for (int t = 0; t < 9; t++) {
          // visiting order: centre, the four face neighbours, the four diagonal rows
          constexpr int oy[9] = {0, -1, 1, 0, 0, -1, 1, -1, 1};
          constexpr int oz[9] = {0, 0, 0, -1, 1, -1, -1, 1, 1};
          const float lbyz = gys[oy[t] + 1] + gzs[oz[t] + 1];
          if (lbyz * s2 > best) continue;  // also true for rows outside the grid (inf) once best is finite ...
          const int yy = cy + oy[t], zz = cz + oz[t];
          if (yy < 0 || yy >= gy || zz < 0 || zz >= gz) continue;  // ... and this covers best == inf
#ifdef MVP_GRID_NOXCLIP
          const int x0 = xlo, x1 = xhi;
#else
          const int x0 = (gxs[0] + lbyz) * s2 > best ? cx : xlo;
          const int x1 = (gxs[2] + lbyz) * s2 > best ? cx : xhi;
#endif
          const int base = (zz * gy + yy) * gx;
          scan(__ldg(start + base + x0), __ldg(start + base + x1 + 1));
          if (budget < 0) break;
        }
      } else {
        for (int dz = -r; dz <= r && budget >= 0; dz++) {
          const int zz = cz + dz;
          if (zz < 0 || zz >= gz) continue;
          const float gzz = gap(uz, zz, slz);
          if (gzz * gzz * s2 > best) continue;
          for (int dy = -r; dy <= r; dy++) {
            const int yy = cy + dy;
            if (yy < 0 || yy >= gy) continue;
            const float gyy = gap(uy, yy, sly);
            const float lbyz = fmaf(gyy, gyy, gzz * gzz);
            if (lbyz * s2 > best) continue;
            if (dz == -r || dz == r || dy == -r || dy == r) {
              row(cx - r, cx + r, yy, zz, lbyz);
            } else {  // interior row: only its two new end cells
              row(cx - r, cx - r, yy, zz, lbyz);
              row(cx + r, cx + r, yy, zz, lbyz);
            }
          }
        }
      }
      if (budget < 0) break;
      // everything outside the cube of radius r is at least `ext` cells away
      float ext = inf;
      if (cx - r > 0) ext = fminf(ext, ux - (float)(cx - r) - slx);
      if (cx + r + 1 < gx) ext = fminf(ext, (float)(cx + r + 1) - ux - slx);
      if (cy - r > 0) ext = fminf(ext, uy - (float)(cy - r) - sly);
      if (cy + r + 1 < gy) ext = fminf(ext, (float)(cy + r + 1) - uy - sly);
      if (cz - r > 0) ext = fminf(ext, uz - (float)(cz - r) - slz);
      if (cz + r + 1 < gz) ext = fminf(ext, (float)(cz + r + 1) - uz - slz);
      ext = fmaxf(ext, 0.f);
      if (ext == inf || best < ext * ext * s2) done = true;
    }
  }
  if (done) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (!kRT || k >= K - kk) {  // real entries follow the phantoms
        const int o = kRT ? k - (K - kk) : k;
        dist[(size_t)orig * ko + o] = bd[k];
        // a slot nothing filled (every remaining candidate had a NaN distance) keeps distance +inf and gets a
        // valid, distinct index: callers index with the result
        idx[(size_t)orig * ko + o] = (kRT && bk[k] == 0x7fffffff) ? o : bk[k];
      }
    }
    if constexpr (K == 3 && !kRT) {  // three_nn: dist2 doubles as the (b, n, 3) weights of three_nn_upsampling
      if (dist2 != nullptr) three_nn_weights_of(bd[0], bd[1], bd[2], dist2 + ((size_t)cloud * nq + orig) * 3);
    }
  } else {
    if (K == 1 && !kRT) {  // what was found before giving up: the completion pass (chamfer_rest.cu) starts from this bound
      dist[orig] = bd[0];
      idx[orig] = bk[0];
    }
    // append to the left-over list of (direction, cloud): one atomic per group of lanes sharing the list
    const int li = dir * b + cloud;
    const unsigned peers = __match_any_sync(__activemask(), li);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    int pos = 0;
    if (lane == leader) pos = atomicAdd(W.count + li, __popc(peers));
    pos = __shfl_sync(peers, pos, leader) + __popc(peers & ((1u << lane) - 1u));
    ((dir ? W.list[1] : W.list[0]) + (size_t)cloud * nq)[pos] = orig;
  }
  }  // t < total

}

// chamfer_rest.cu: the completion pass over the left-over lists
int chamfer_rest_launch(int b, int n, int m, const GridWs &W, const float *xyz1, const float *xyz2, float *dist1,
                        float *dist2, int *idx1, int *idx2, cudaStream_t s);

bool chamfer_grid_supported(int b, int n, int m) {
  return b > 0 && b <= 65535 && n >= 512 && m >= 512 && n <= (1 << 20) && m <= (1 << 20) &&
         (long long)b * ((long long)n + m) < (1LL << 31);
}

size_t chamfer_grid_workspace_bytes(int b, int n, int m) { return grid_plan(b, n, m, nullptr, nullptr); }

int chamfer_grid_launch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1, float *dist2,
                        int *idx1, int *idx2, void *ws, size_t ws_bytes, cudaStream_t s) {
  GridWs W;
  if (ws_bytes < grid_plan(b, n, m, ws, &W)) return MVP_ERR_WORKSPACE;
  {
    const int rc = grid_build_launch(b, n, m, xyz1, xyz2, W, s);
    if (rc) return rc;
  }
  const long long total = (long long)b * ((long long)n + m);
  static const int dyn = [] {  // measuring aid: MVP_GRID_DYNROWS=0 selects the rows-in-lockstep kernel
    const char *e = getenv("MVP_GRID_DYNROWS");
    return e ? atoi(e) : 1;
  }();
  static const int pdl = [] {  // measuring aid: MVP_PDL=0 launches the chain the ordinary way
    const char *e = getenv("MVP_PDL");
    return e ? atoi(e) : 1;
  }();
  static const int pool = [] {  // measuring aid: MVP_GRID_POOL=0 keeps every query's rows on its own lane
    const char *e = getenv("MVP_GRID_POOL");
    return e ? atoi(e) : 1;
  }();
  if (dyn && pool) {
    const cudaError_t e = launch_pdl(chamfer_grid_query_kernel<1, false, true, true>,
                                     dim3((unsigned)((total + kGridQThreads - 1) / kGridQThreads)), dim3(kGridQThreads), 0, s, b, n,
                                     m, W, dist1, dist2, idx1, idx2, 1);
    if (e != cudaSuccess) return (int)e;
  } else if (dyn && pdl) {
    const cudaError_t e = launch_pdl(chamfer_grid_query_kernel<1, false, true>,
                                     dim3((unsigned)((total + kGridQThreads - 1) / kGridQThreads)), dim3(kGridQThreads), 0, s, b, n,
                                     m, W, dist1, dist2, idx1, idx2, 1);
    if (e != cudaSuccess) return (int)e;
  } else if (dyn)
    chamfer_grid_query_kernel<1, false, true><<<(unsigned)((total + kGridQThreads - 1) / kGridQThreads), kGridQThreads, 0, s>>>(
        b, n, m, W, dist1, dist2, idx1, idx2, 1);
  else
    chamfer_grid_query_kernel<1><<<(unsigned)((total + kGridQThreads - 1) / kGridQThreads), kGridQThreads, 0, s>>>(
        b, n, m, W, dist1, dist2, idx1, idx2, 1);
  count_launch(2);
  const int rc = launch_status();
  if (rc) return rc;
  // what the search did not finish (normally nothing: the kernel leaves at once)
  return chamfer_rest_launch(b, n, m, W, xyz1, xyz2, dist1, dist2, idx1, idx2, s);
}

// ---- three_nn through the same grid (pointnet2.cu: mvp_three_nn_ws) --------------------------------------------------
// pointnet2.cu
int three_nn_rest_launch(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                         float *weight, const int *list, const int *count, cudaStream_t s);

bool three_nn_grid_supported(int b, int n, int m) {
  return b > 0 && b <= 65535 && n >= 256 && m >= 256 && n <= (1 << 20) && m <= (1 << 20) &&
         (long long)b * ((long long)n + m) < (1LL << 31);
}

size_t three_nn_grid_workspace_bytes(int b, int n, int m) { return grid_plan(b, n, m, nullptr, nullptr); }

// unknown (b,n,3) against known (b,m,3): grid over `known`, queries in the sorted order of `unknown`; queries that do
// not finish within the ring / candidate budget are finished by the tiled brute-force kernel through a list.
int three_nn_grid_launch(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                         float *weight, void *ws, size_t ws_bytes, cudaStream_t s) {
  GridWs W;
  if (ws_bytes < grid_plan(b, n, m, ws, &W)) return MVP_ERR_WORKSPACE;
  {
    const int rc = grid_build_launch(b, n, m, unknown, known, W, s);
    if (rc) return rc;
  }
  const long long total = (long long)b * n;
  chamfer_grid_query_kernel<3><<<(unsigned)((total + kGridQThreads - 1) / kGridQThreads), kGridQThreads, 0, s>>>(
      b, n, m, W, dist2, weight, idx, nullptr, 3);  // (the second distance output carries the weights for K = 3)
  count_launch(2);
  int rc = launch_status();
  if (rc) return rc;
  return three_nn_rest_launch(b, n, m, unknown, known, dist2, idx, weight, W.list[0], W.count, s);
}


// ---- k nearest neighbours of 3-D points (mvp_knn_points) ---------------------------------------------------------------
// Exhaustive top-k: one thread per query, the searched cloud broadcast from shared memory; the same ordered list and
// the same (distance, index) order as the grid search.  kRest: the queries are the count[b] entries of list + b*n
// (what the grid search did not finish); CTAs beyond the list length leave at once.  Also the whole search for
// clouds too small for a grid to pay.
constexpr int kTopkTile = 1024;
template <int K, bool kRest>
__global__ void __launch_bounds__(128)
topk_nn_kernel(int n, int m, int kk, const float *__restrict__ queries, const float *__restrict__ cloud,
               float *__restrict__ dist2, int *__restrict__ idx, const int *__restrict__ list,
               const int *__restrict__ count) {
  __shared__ float4 tile[kTopkTile];
  const int b = blockIdx.y;
  const int nq = kRest ? __ldg(count + b) : n;
  if (blockIdx.x * 128 >= nq) return;
  int p = blockIdx.x * 128 + threadIdx.x;
  const bool active = p < nq;
  if (kRest && active) p = __ldg(list + (size_t)b * n + p);
  const float *C = cloud + (size_t)b * m * 3;
  float qx = 0, qy = 0, qz = 0;
  if (active) {
    const float *q = queries + ((size_t)b * n + p) * 3;
    qx = __ldg(q + 0), qy = __ldg(q + 1), qz = __ldg(q + 2);
  }
  const float inf = __int_as_float(0x7f800000);
  float bd[K];
  int bk[K];
#pragma unroll
  for (int k = 0; k < K; k++) {  // K - kk phantoms at -inf in front: the kk-th real neighbour is slot K - 1
    bd[k] = k < K - kk ? -inf : inf;
    bk[k] = k < K - kk ? -1 : 0x7fffffff;
  }
  for (int j0 = 0; j0 < m; j0 += kTopkTile) {
    const int cnt = min(kTopkTile, m - j0);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += 128) {
      const float *c = C + (size_t)(j0 + j) * 3;
      tile[j] = make_float4(__ldg(c + 0), __ldg(c + 1), __ldg(c + 2), 0.f);
    }
    __syncthreads();
    if (!active) continue;
    for (int j = 0; j < cnt; j++) {
      const float4 c = tile[j];
      const float d = sqdist(c.x - qx, c.y - qy, c.z - qz);
      // index order: an equal distance never precedes an earlier one; NaN distances are never kept
      if (d < bd[K - 1] || (d == bd[K - 1] && j0 + j < bk[K - 1])) topk_insert<K>(bd, bk, d, j0 + j);
    }
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (k >= K - kk) {
        dist2[((size_t)b * n + p) * kk + k - (K - kk)] = bd[k];
        // unfilled slot (NaN distances are never kept): distance stays +inf, the index a valid distinct one
        idx[((size_t)b * n + p) * kk + k - (K - kk)] = bk[k] == 0x7fffffff ? k - (K - kk) : bk[k];
      }
    }
  }
}

bool knn_points_grid_supported(int b, int n, int m, int k) {
  return k <= 32 && b > 0 && b <= 65535 && n >= 256 && m >= 256 && n <= (1 << 20) && m <= (1 << 20) &&
         (long long)b * ((long long)n + m) < (1LL << 31);
}

size_t knn_points_workspace_bytes(int b, int n, int m, int k) {
  return knn_points_grid_supported(b, n, m, k) ? grid_plan(b, n, m, nullptr, nullptr) : 16;
}

template <int K, bool kGrid = true>
static int knn_points_launch_k(int b, int n, int m, int k, const float *queries, const float *cloud, float *dist2,
                               int *idx, void *ws, size_t ws_bytes, cudaStream_t s) {
  if constexpr (kGrid)
  if (knn_points_grid_supported(b, n, m, k) && ws && ws_bytes >= grid_plan(b, n, m, nullptr, nullptr)) {
    GridWs W;
    grid_plan(b, n, m, ws, &W);
      {
      const int rc = grid_build_launch(b, n, m, queries, cloud, W, s);
      if (rc) return rc;
    }
    const long long total = (long long)b * n;
    chamfer_grid_query_kernel<K, true><<<(unsigned)((total + kGridQThreads - 1) / kGridQThreads), kGridQThreads, 0, s>>>(
        b, n, m, W, dist2, nullptr, idx, nullptr, k);
    topk_nn_kernel<K, true><<<dim3((n + 127) / 128, b), 128, 0, s>>>(n, m, k, queries, cloud, dist2, idx, W.list[0],
                                                                      W.count);
    count_launch(3);
    return launch_status();
  }
  for (int b0 = 0; b0 < b; b0 += 65535) {
    const int bb = std::min(65535, b - b0);
    topk_nn_kernel<K, false><<<dim3((n + 127) / 128, bb), 128, 0, s>>>(
        n, m, k, queries + (size_t)b0 * n * 3, cloud + (size_t)b0 * m * 3, dist2 + (size_t)b0 * n * k,
        idx + (size_t)b0 * n * k, nullptr, nullptr);
    count_launch();
  }
  return launch_status();
}

int knn_points_launch(int b, int n, int m, int k, const float *queries, const float *cloud, float *dist2, int *idx,
                      void *ws, size_t ws_bytes, cudaStream_t s) {
  if (k <= 4) return knn_points_launch_k<4>(b, n, m, k, queries, cloud, dist2, idx, ws, ws_bytes, s);
  if (k <= 8) return knn_points_launch_k<8>(b, n, m, k, queries, cloud, dist2, idx, ws, ws_bytes, s);
  if (k <= 12) return knn_points_launch_k<12>(b, n, m, k, queries, cloud, dist2, idx, ws, ws_bytes, s);
  if (k <= 16) return knn_points_launch_k<16>(b, n, m, k, queries, cloud, dist2, idx, ws, ws_bytes, s);
  if (k <= 24) return knn_points_launch_k<24>(b, n, m, k, queries, cloud, dist2, idx, ws, ws_bytes, s);
  if (k <= 32) return knn_points_launch_k<32>(b, n, m, k, queries, cloud, dist2, idx, ws, ws_bytes, s);
  return knn_points_launch_k<64, false>(b, n, m, k, queries, cloud, dist2, idx, nullptr, 0, s);  // exhaustive only
}

}  // namespace mvp

// k nearest points of `cloud` (b,m,3) for every query (b,n,3): squared distances and indices (b,n,k) in ascending
// (distance, index) order; k <= min(m, 64).  See include/mvp_ops.h.
MVP_API size_t mvp_knn_points_workspace_bytes(int b, int n, int m, int k) {
  if (b < 0 || n < 0 || m < 0 || k < 1) return 16;
  return mvp::knn_points_workspace_bytes(b, n, m, k);
}

MVP_API int mvp_knn_points(int b, int n, int m, int k, const float *queries, const float *cloud, float *dist2, int *idx,
                           void *workspace, size_t workspace_bytes, mvp_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || k < 1 || k > 64 || k > m) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0) return MVP_OK;
  if (!queries || !cloud || !dist2 || !idx) return MVP_ERR_INVALID_ARGUMENT;
  return mvp::knn_points_launch(b, n, m, k, queries, cloud, dist2, idx, workspace, workspace_bytes, (cudaStream_t)stream);
}
