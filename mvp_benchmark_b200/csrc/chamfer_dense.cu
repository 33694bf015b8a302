// Chamfer distance forward, warp-cooperative grid path for sm_100a — the default for clouds of 512..16384 points.
//
// Same outputs, bit for bit, as the brute-force kernels (chamfer.cu / chamfer_fused.cu) and therefore as the
// reference's NmDistanceKernel (utils/metrics/CD/chamfer3D/chamfer3D.cu:12-134): every evaluated candidate uses the
// reference's contraction  d = fma(dz,dz, fma(dx,dx, dy*dy)),  equal minima resolve to the lowest index (:36,:126), and
// a candidate is skipped only when a conservative lower bound of its distance is strictly above a distance found.
//
// Where chamfer_grid.cu gives every query its own thread (a private walk over 27 cells: 13 of 32 lanes active per
// issued instruction, ncu r1), this path makes the WARP the unit of work and keeps its lanes full:
//
//   build   one cluster of 2 or 4 CTAs per cloud pair (chamfer_dense_build_kernel).  Each cloud is counting-sorted
//           TWICE in the same passes: by the cells of its OWN uniform grid (the target structure: float4 (x, y, z,
//           original index) + cell starts, x the fastest cell coordinate) and by the 2x2x2-cell BLOCKS of the OTHER
//           cloud's grid (the query order).  Both headers are known to every CTA of the cluster after one exchange of
//           partial bounding boxes through distributed shared memory.
//   query   chamfer_dense_query_kernel: a warp owns 32 consecutive queries of the block-sorted order.  For each
//           block they fall in, the 4x4x4 cells around the block (the block and a one-cell halo: 16 contiguous
//           x-runs of the sorted target array) are brought into shared memory by 16 bulk copies (cp.async.bulk +
//           mbarrier, one per lane, SASS UBLKCP) and from there into registers, LANES OVER CANDIDATES; the block's
//           queries are then broadcast one by one: packed fp32x2 distances (3 issue slots per 2 candidates), FMNMX3,
//           one redux.sync.min over the bit patterns, one ballot to name the winning lane.  Which of that lane's
//           candidates won — and the lowest original index among equal minima — is resolved afterwards, lanes over
//           QUERIES, by re-evaluating only the winning lane's few candidates.  A query is finished when its best
//           distance is below the distance to everything outside the halo (same conservative bound as chamfer_grid.cu).
//   rest    queries that are not (far outside the other cloud, sparse surroundings, degenerate grids) keep their
//           provisional result as a bound and are completed by chamfer_dense_rest_kernel: one thread per query, every
//           row of cells of the target grid either skipped by its lower bound or scanned over the x-range the bound
//           allows.  Normally a fraction of a percent of the queries; the kernel leaves at once when there are none.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"
#include "grid.cuh"
#include "sm100.cuh"

namespace mvp {

namespace cg = cooperative_groups;

#ifndef MVP_DENSE_PPC_X2
#define MVP_DENSE_PPC_X2 3            // target points per cell, times two (3 = 1.5 points: ~100 candidates per halo)
#endif
constexpr int kDenseMinPts = 512;
constexpr int kDenseMaxPts = 16384;   // per cloud: two build CTAs of 8192 points per side
constexpr int kDbThreads = 1024;      // build CTA
constexpr int kDbPts = 8;             // points per build thread
constexpr int kDbPerO = 12;           // own-grid cells per build thread in the scan (<= 12288 cells)
constexpr int kDbPerB = 8;            // other-grid blocks per build thread in the scan (<= 8192 blocks)
constexpr int kDbSplitPts = 6144;     // clouds from this size up are split over two CTAs per side

constexpr int kDqWarps = 4;           // query CTA: four independent warps
constexpr int kDqThreads = 32 * kDqWarps;
constexpr int kDqTile = 256;          // candidate slots per shared-memory tile: 8 per lane
#ifndef MVP_DENSE_TMA
#define MVP_DENSE_TMA 0               // 1: stage the halo rows with one cp.async.bulk per row (measured 2x slower: ~100-byte copies)
#endif
#ifndef MVP_DENSE_QMINB
#define MVP_DENSE_QMINB 8             // resident query CTAs per SM ptxas is asked for (64 registers)
#endif

struct DenseWs {  // carved out of the caller's workspace by dense_plan()
  GridHdr *hdr;         // [2][b]   pad[0..2] = blocks per axis
  int *count;           // [2][b]   left-over list lengths
  int *start[2];        // [b][cap_side + 1]
  float4 *sorted[2];    // [b][n] / [b][m]   by cell of the cloud's own grid
  float4 *qsorted[2];   // [b][n] / [b][m]   by block of the OTHER cloud's grid
  int *list[2];         // [b][n] / [b][m]   left-over queries (original indices)
  int cap[2];
};

static int dense_cap(int npts) {
  const long long c = ((long long)npts * 2) / MVP_DENSE_PPC_X2;
  return (int)std::min<long long>(std::max<long long>(c, 8), (long long)kDbPerO * kDbThreads);
}

static size_t dense_plan(int b, int n, int m, void *base, DenseWs *w) {
  size_t off = 0;
  unsigned char *p = reinterpret_cast<unsigned char *>(base);
  auto take = [&](size_t bytes) {
    unsigned char *r = p ? p + off : nullptr;
    off += (bytes + 15) & ~(size_t)15;
    return r;
  };
  DenseWs t;
  t.cap[0] = dense_cap(n);
  t.cap[1] = dense_cap(m);
  t.hdr = reinterpret_cast<GridHdr *>(take(sizeof(GridHdr) * 2 * (size_t)b));
  t.count = reinterpret_cast<int *>(take(sizeof(int) * 2 * (size_t)b));
  t.start[0] = reinterpret_cast<int *>(take(sizeof(int) * (size_t)b * (t.cap[0] + 1)));
  t.start[1] = reinterpret_cast<int *>(take(sizeof(int) * (size_t)b * (t.cap[1] + 1)));
  t.sorted[0] = reinterpret_cast<float4 *>(take(sizeof(float4) * (size_t)b * n));
  t.sorted[1] = reinterpret_cast<float4 *>(take(sizeof(float4) * (size_t)b * m));
  t.qsorted[0] = reinterpret_cast<float4 *>(take(sizeof(float4) * (size_t)b * n));
  t.qsorted[1] = reinterpret_cast<float4 *>(take(sizeof(float4) * (size_t)b * m));
  t.list[0] = reinterpret_cast<int *>(take(sizeof(int) * (size_t)b * n));
  t.list[1] = reinterpret_cast<int *>(take(sizeof(int) * (size_t)b * m));
  if (w) *w = t;
  return off;
}

// cell and block of a point in the grid `h`; the SAME expression in the build and in the query kernel
struct CellPos {
  float ux, uy, uz;
  int cx, cy, cz;
};
__device__ __forceinline__ CellPos cell_pos(const GridHdr &h, float x, float y, float z) {
  CellPos p;
  p.ux = (x - h.lo[0]) * h.inv_s;
  p.uy = (y - h.lo[1]) * h.inv_s;
  p.uz = (z - h.lo[2]) * h.inv_s;
  p.cx = cell_coord(p.ux, h.g[0]);
  p.cy = cell_coord(p.uy, h.g[1]);
  p.cz = cell_coord(p.uz, h.g[2]);
  return p;
}

// ------------------------------------------------------------------------------------------------ build
// Cluster of 2 * PARTS CTAs per cloud pair: ranks [0, PARTS) sort xyz1, ranks [PARTS, 2 PARTS) sort xyz2; with
// PARTS == 2 each CTA owns half of its cloud's points and the partial histograms meet through distributed shared
// memory (the scheme of chamfer_grid_build2_kernel), CTA 1's cursor of a cell starting behind CTA 0's points.
template <int PARTS>
__global__ void __launch_bounds__(kDbThreads, 1)
chamfer_dense_build_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                           DenseWs W) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) int db_smem[];
  int *hist_o = db_smem;                                   // [kDbPerO * kDbThreads] cells of the own grid
  int *hist_b = hist_o + kDbPerO * kDbThreads;             // [kDbPerB * kDbThreads] blocks of the other grid
  unsigned *code_o = reinterpret_cast<unsigned *>(hist_b + kDbPerB * kDbThreads);  // [kDbPts * kDbThreads]
  unsigned *code_b = code_o + kDbPts * kDbThreads;
  __shared__ float s_red[6][32];
  __shared__ int s_fin[32];
  __shared__ int s_warp[2][32];
  __shared__ float s_part[8];  // this CTA's box (lo xyz, hi xyz) and finiteness flag, read by the whole cluster
  __shared__ GridHdr s_hdr[2];
  const int crank = (int)cluster.block_rank();
  const int cloud = blockIdx.x / (2 * PARTS);
  const int side = crank / PARTS, part = crank % PARTS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = side ? m : n;
  const float *P = (side ? xyz2 : xyz1) + (size_t)cloud * np * 3;
  const int half0 = PARTS == 2 ? (np + 1) / 2 : np;
  const int first = part ? half0 : 0, mine = part ? np - half0 : half0;  // this CTA's points: [first, first + mine)
  const float inf = __int_as_float(0x7f800000);

  // ---- pass 1: bounding box and finiteness of this CTA's points
  float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  int fin = 1;
#pragma unroll
  for (int u = 0; u < kDbPts; u++) {
    const int j = tid + u * kDbThreads;
    if (j < mine) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const float v = __ldg(P + (size_t)(first + j) * 3 + a);
        lo[a] = fminf(lo[a], v);
        hi[a] = fmaxf(hi[a], v);
        fin &= (fabsf(v) <= 3.0e38f) ? 1 : 0;  // false for NaN and +-inf
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
    for (int w = 1; w < kDbThreads / 32; w++) {
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
  cluster.sync();
  // ---- both headers, identically in every CTA of the cluster (min / max commute): warp 0 derives side 0's, warp 1
  // side 1's
  if (lane == 0 && warp < 2) {
    const int sd = warp;
    float blo[3] = {inf, inf, inf}, bhi[3] = {-inf, -inf, -inf};
    int bfin = 1;
    for (int pr = 0; pr < PARTS; pr++) {
      const float *q = cluster.map_shared_rank(s_part, sd * PARTS + pr);
#pragma unroll
      for (int a = 0; a < 3; a++) {
        blo[a] = fminf(blo[a], q[a]);
        bhi[a] = fmaxf(bhi[a], q[3 + a]);
      }
      bfin &= __float_as_int(q[6]);
    }
    GridHdr h = grid_header(blo, bhi, bfin, sd ? W.cap[1] : W.cap[0]);
#pragma unroll
    for (int a = 0; a < 3; a++) h.pad[a] = (h.g[a] + 1) >> 1;  // 2x2x2-cell blocks per axis
    s_hdr[sd] = h;
    if (sd == side && part == 0) {
      W.hdr[side * b + cloud] = h;
      W.count[side * b + cloud] = 0;
    }
  }
  __syncthreads();
  const GridHdr ho = s_hdr[side], hb = s_hdr[1 - side];  // own grid (cells), other grid (blocks)
  const int ncell = ho.ncell;
  const int nbx = hb.pad[0], nby = hb.pad[1];
  {
    int4 *z = reinterpret_cast<int4 *>(db_smem);
    for (int c = tid; c < (kDbPerO + kDbPerB) * kDbThreads / 4; c += kDbThreads) z[c] = make_int4(0, 0, 0, 0);
  }
  __syncthreads();

  // ---- pass 2: both histograms; every point keeps (bin, arrival rank) of both sorts for the scatter
#pragma unroll
  for (int u = 0; u < kDbPts; u++) {
    const int j = tid + u * kDbThreads;
    if (j < mine) {
      const float *q = P + (size_t)(first + j) * 3;
      const float x = __ldg(q), y = __ldg(q + 1), z = __ldg(q + 2);
      int c = 0, k = 0;
      if (ho.valid) {
        const CellPos p = cell_pos(ho, x, y, z);
        c = (p.cz * ho.g[1] + p.cy) * ho.g[0] + p.cx;
      }
      if (hb.valid) {
        const CellPos p = cell_pos(hb, x, y, z);
        k = ((p.cz >> 1) * nby + (p.cy >> 1)) * nbx + (p.cx >> 1);
      }
      code_o[j] = ((unsigned)c << 15) | (unsigned)atomicAdd(&hist_o[c], 1);  // c < 2^15 cells, rank < 2^13
      code_b[j] = ((unsigned)k << 15) | (unsigned)atomicAdd(&hist_b[k], 1);
    }
  }
  cluster.sync();  // the partial histograms of the whole cluster are complete

  // ---- exclusive scans over (own + peer) counts.  The counts sit in registers across the barrier after which this
  // CTA overwrites its histograms with its fill cursors.
  const int peer = side * PARTS + (part ^ 1);
  int own_o[kDbPerO], oth_o[kDbPerO], own_b[kDbPerB], oth_b[kDbPerB];
  int sum_o = 0, sum_b = 0;
  {
    const int4 *a4 = reinterpret_cast<const int4 *>(hist_o) + tid * (kDbPerO / 4);
#pragma unroll
    for (int e = 0; e < kDbPerO / 4; e++) {
      const int4 v = a4[e];
      own_o[4 * e] = v.x, own_o[4 * e + 1] = v.y, own_o[4 * e + 2] = v.z, own_o[4 * e + 3] = v.w;
    }
    const int4 *b4 = reinterpret_cast<const int4 *>(hist_b) + tid * (kDbPerB / 4);
#pragma unroll
    for (int e = 0; e < kDbPerB / 4; e++) {
      const int4 v = b4[e];
      own_b[4 * e] = v.x, own_b[4 * e + 1] = v.y, own_b[4 * e + 2] = v.z, own_b[4 * e + 3] = v.w;
    }
    if (PARTS == 2) {
      const int4 *pa4 = reinterpret_cast<const int4 *>(cluster.map_shared_rank(hist_o, peer)) + tid * (kDbPerO / 4);
#pragma unroll
      for (int e = 0; e < kDbPerO / 4; e++) {
        const int4 v = pa4[e];
        oth_o[4 * e] = v.x, oth_o[4 * e + 1] = v.y, oth_o[4 * e + 2] = v.z, oth_o[4 * e + 3] = v.w;
      }
      const int4 *pb4 = reinterpret_cast<const int4 *>(cluster.map_shared_rank(hist_b, peer)) + tid * (kDbPerB / 4);
#pragma unroll
      for (int e = 0; e < kDbPerB / 4; e++) {
        const int4 v = pb4[e];
        oth_b[4 * e] = v.x, oth_b[4 * e + 1] = v.y, oth_b[4 * e + 2] = v.z, oth_b[4 * e + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < kDbPerO; e++) oth_o[e] = 0;
#pragma unroll
      for (int e = 0; e < kDbPerB; e++) oth_b[e] = 0;
    }
#pragma unroll
    for (int e = 0; e < kDbPerO; e++) sum_o += own_o[e] + oth_o[e];  // bins past the last one hold zeros
#pragma unroll
    for (int e = 0; e < kDbPerB; e++) sum_b += own_b[e] + oth_b[e];
  }
  int incl_o = sum_o, incl_b = sum_b;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int vo = __shfl_up_sync(0xffffffffu, incl_o, off);
    const int vb = __shfl_up_sync(0xffffffffu, incl_b, off);
    if (lane >= off) incl_o += vo, incl_b += vb;
  }
  if (lane == 31) s_warp[0][warp] = incl_o, s_warp[1][warp] = incl_b;
  cluster.sync();  // (also a CTA barrier) the peer has read this CTA's histograms: they may be overwritten now
  if (warp < 2) {
    int v = s_warp[warp][lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, off);
      if (lane >= off) v += u;
    }
    s_warp[warp][lane] = v;
  }
  __syncthreads();
  {
    int run = incl_o - sum_o + (warp ? s_warp[0][warp - 1] : 0);
    int4 *a4 = reinterpret_cast<int4 *>(hist_o) + tid * (kDbPerO / 4);
    int cur[kDbPerO];
#pragma unroll
    for (int e = 0; e < kDbPerO; e++) {
      cur[e] = run + (part ? oth_o[e] : 0);  // CTA 1 fills a bin behind CTA 0's points
      run += own_o[e] + oth_o[e];
    }
#pragma unroll
    for (int e = 0; e < kDbPerO / 4; e++) a4[e] = make_int4(cur[4 * e], cur[4 * e + 1], cur[4 * e + 2], cur[4 * e + 3]);
  }
  {
    int run = incl_b - sum_b + (warp ? s_warp[1][warp - 1] : 0);
    int4 *b4 = reinterpret_cast<int4 *>(hist_b) + tid * (kDbPerB / 4);
    int cur[kDbPerB];
#pragma unroll
    for (int e = 0; e < kDbPerB; e++) {
      cur[e] = run + (part ? oth_b[e] : 0);
      run += own_b[e] + oth_b[e];
    }
#pragma unroll
    for (int e = 0; e < kDbPerB / 4; e++) b4[e] = make_int4(cur[4 * e], cur[4 * e + 1], cur[4 * e + 2], cur[4 * e + 3]);
  }
  __syncthreads();
  if (part == 0) {  // CTA 0's cursors are the cell starts: coalesced copy
    int *start = (side ? W.start[1] : W.start[0]) + (size_t)cloud * ((side ? W.cap[1] : W.cap[0]) + 1);
    for (int c = tid; c < ncell; c += kDbThreads) start[c] = hist_o[c];
    if (tid == 0) start[ncell] = np;
  }

  // ---- pass 3: scatter into both orders (the order inside a bin is arbitrary; the query's tie rule is explicit)
  float4 *S = (side ? W.sorted[1] : W.sorted[0]) + (size_t)cloud * np;
  float4 *Q = (side ? W.qsorted[1] : W.qsorted[0]) + (size_t)cloud * np;
#pragma unroll
  for (int u = 0; u < kDbPts; u++) {
    const int j = tid + u * kDbThreads;
    if (j < mine) {
      const float *q = P + (size_t)(first + j) * 3;
      const float4 v = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __int_as_float(first + j));
      const unsigned co = code_o[j], cb = code_b[j];
      S[hist_o[co >> 15] + (int)(co & 0x7fffu)] = v;
      Q[hist_b[cb >> 15] + (int)(cb & 0x7fffu)] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ query
struct DqSmem {
  float4 cand[kDqWarps][kDqTile];
  float4 qry[kDqWarps][32];
  int row_off[kDqWarps][16];  // first slot of each halo row in the block's candidate list
  int row_pos[kDqWarps][16];  // position of that slot in the sorted target array
#if MVP_DENSE_TMA
  uint64_t bar[kDqWarps];
#endif
};

// One tile of candidates (navail <= kDqTile slots of sC, slot = lane + 32 k) against the queries of the lanes in
// `grp`: per query the minimum over the tile as a bit pattern and the lane that holds it (bit 8: several lanes do).
template <int K2>
__device__ __forceinline__ void dq_tile(const float4 *sC, const float4 *sQ, unsigned grp, int lane, int navail,
                                        uint32_t &tbest, int &twin) {
  const float nanv = __int_as_float(0x7fffffff);  // an absent slot: its distance is NaN, which a minimum ignores
  u64 X[K2], Y[K2], Z[K2];
#pragma unroll
  for (int h = 0; h < K2; h++) {
    const int s0 = lane + 64 * h, s1 = s0 + 32;
    float4 c0 = make_float4(nanv, nanv, nanv, 0.f), c1 = c0;
    if (s0 < navail) c0 = sC[s0];
    if (s1 < navail) c1 = sC[s1];
    X[h] = pack2(c0.x, c1.x), Y[h] = pack2(c0.y, c1.y), Z[h] = pack2(c0.z, c1.z);
  }
  unsigned mq = grp;
  while (mq) {
    const int l = __ffs(mq) - 1;
    mq &= mq - 1;
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
    const unsigned bal = __ballot_sync(0xffffffffu, mb == mw);
    if (lane == l) {
      tbest = mw;
      twin = (__ffs(bal) - 1) | (__popc(bal) > 1 ? 256 : 0);
    }
  }
}

__global__ void __launch_bounds__(kDqThreads, MVP_DENSE_QMINB)
chamfer_dense_query_kernel(int b, int n, int m, DenseWs W, float *__restrict__ dist1, float *__restrict__ dist2,
                           int *__restrict__ idx1, int *__restrict__ idx2) {
  __shared__ __align__(16) DqSmem S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 *sC = S.cand[warp];
  float4 *sQ = S.qry[warp];
  int *sOff = S.row_off[warp], *sPos = S.row_pos[warp];
#if MVP_DENSE_TMA
  uint64_t *bar = &S.bar[warp];
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t parity = 0;
#endif

  const int chunks1 = (n + 31) >> 5, chunks2 = (m + 31) >> 5, per_cloud = chunks1 + chunks2;
  const long long wid = (long long)blockIdx.x * kDqWarps + warp;
  if (wid >= (long long)b * per_cloud) return;
  const int cloud = (int)(wid / per_cloud), rem = (int)(wid % per_cloud);
  const int dir = rem >= chunks1 ? 1 : 0;  // 0: points of xyz1 against xyz2's grid; 1: the other way round
  const int chunk = dir ? rem - chunks1 : rem;
  const int nq = dir ? m : n, nt = dir ? n : m, ts = 1 - dir;
  const GridHdr *hp = W.hdr + ts * b + cloud;
  const int4 h0 = __ldg(reinterpret_cast<const int4 *>(hp));      // lo.xyz, inv_s
  const int4 h1 = __ldg(reinterpret_cast<const int4 *>(hp) + 1);  // s, g.xyz
  const int4 h2 = __ldg(reinterpret_cast<const int4 *>(hp) + 2);  // ncell, valid, blocks x, blocks y
  GridHdr h;
  h.lo[0] = __int_as_float(h0.x), h.lo[1] = __int_as_float(h0.y), h.lo[2] = __int_as_float(h0.z);
  h.inv_s = __int_as_float(h0.w);
  h.s = __int_as_float(h1.x);
  h.g[0] = h1.y, h.g[1] = h1.z, h.g[2] = h1.w;
  const int gx = h.g[0], gy = h.g[1], gz = h.g[2];
  const int grid_ok = h2.y;
  const int *start = (ts ? W.start[1] : W.start[0]) + (size_t)cloud * ((ts ? W.cap[1] : W.cap[0]) + 1);
  const float4 *T = (ts ? W.sorted[1] : W.sorted[0]) + (size_t)cloud * nt;
  const float4 *Q = (dir ? W.qsorted[1] : W.qsorted[0]) + (size_t)cloud * nq;
  float *dist = (dir ? dist2 : dist1) + (size_t)cloud * nq;
  int *idx = (dir ? idx2 : idx1) + (size_t)cloud * nq;

  const int qi = chunk * 32 + lane;
  const bool present = qi < nq;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (present) q = __ldg(Q + qi);
  sQ[lane] = q;
  const CellPos p = cell_pos(h, q.x, q.y, q.z);
  const bool finite = fabsf(p.ux) + fabsf(p.uy) + fabsf(p.uz) < 3.0e38f;  // false for NaN / inf
  const bool active = present && finite && grid_ok;
  const int bx = p.cx >> 1, by = p.cy >> 1, bz = p.cz >> 1;
  const int blk = (bz * h2.w + by) * h2.z + bx;
  uint32_t best = 0x7f800000u;  // +inf
  int bidx = 0x7fffffff;
  __syncwarp();

  unsigned pending = __ballot_sync(0xffffffffu, active);
  while (pending) {
    const int leader = __ffs(pending) - 1;
    const int bb = __shfl_sync(0xffffffffu, blk, leader);
    const unsigned grp = __ballot_sync(0xffffffffu, active && blk == bb) & pending;
    pending &= ~grp;
    const int Bx = __shfl_sync(0xffffffffu, bx, leader), By = __shfl_sync(0xffffffffu, by, leader),
              Bz = __shfl_sync(0xffffffffu, bz, leader);
    // the 16 rows of the halo: lane r < 16 owns row (y, z) = (2 By - 1 + (r & 3), 2 Bz - 1 + (r >> 2)), cells
    // [2 Bx - 1, 2 Bx + 2] of it, one contiguous range of the sorted array
    int a = 0, len = 0;
    if (lane < 16) {
      const int yy = 2 * By - 1 + (lane & 3), zz = 2 * Bz - 1 + (lane >> 2);
      if (yy >= 0 && yy < gy && zz >= 0 && zz < gz) {
        const int x0 = max(2 * Bx - 1, 0), x1 = min(2 * Bx + 2, gx - 1);
        const int base = (zz * gy + yy) * gx;
        a = __ldg(start + base + x0);
        len = __ldg(start + base + x1 + 1) - a;
      }
    }
    int incl = len;
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 15);
    const int off_r = incl - len;
    const bool mine = (grp >> lane) & 1u;
    for (int t0 = 0; t0 < total; t0 += kDqTile) {
      const int navail = min(total - t0, kDqTile);
      __syncwarp();  // every lane is done with the previous contents of the tile
#if MVP_DENSE_TMA
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)navail * 16u);
      {
        const int lo = max(off_r, t0), hi = min(off_r + len, t0 + kDqTile);
        if (hi > lo) tma_load_1d(sC + (lo - t0), T + a + (lo - off_r), (uint32_t)(hi - lo) * 16u, bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1u;
#else
      // slot s of the list lies in the last row whose first slot is <= s (empty rows share their successor's first
      // slot and are never chosen): binary search of the 16 offsets, then one 16-byte load per slot, all slots of a
      // lane in flight together
      if (t0 == 0) {
        if (lane < 16) sOff[lane] = off_r, sPos[lane] = a;
        __syncwarp();
      }
#pragma unroll 1
      for (int k0 = 0; k0 < kDqTile / 32 && 32 * k0 < navail; k0 += 4) {  // four slots per lane at a time
        float4 cv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int sl = lane + 32 * (k0 + k);
          if (sl < navail) {
            const int g = t0 + sl;
            int r = sOff[8] <= g ? 8 : 0;
            r += sOff[r + 4] <= g ? 4 : 0;
            r += sOff[r + 2] <= g ? 2 : 0;
            r += sOff[r + 1] <= g ? 1 : 0;
            cv[k] = __ldg(T + sPos[r] + (g - sOff[r]));
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (lane + 32 * (k0 + k) < navail) sC[lane + 32 * (k0 + k)] = cv[k];
      }
      __syncwarp();
#endif
      uint32_t tbest = 0x7fffffffu;
      int twin = 0;
      const int k2 = (navail + 63) >> 6;
      if (k2 == 1) dq_tile<1>(sC, sQ, grp, lane, navail, tbest, twin);
      else if (k2 == 2) dq_tile<2>(sC, sQ, grp, lane, navail, tbest, twin);
      else if (k2 == 3) dq_tile<3>(sC, sQ, grp, lane, navail, tbest, twin);
      else dq_tile<4>(sC, sQ, grp, lane, navail, tbest, twin);
      // which candidate it was, lanes over queries: the winning lane's slots (or, when several lanes hold the
      // minimum, every slot of the tile) re-evaluated with the scalar form of the same contraction; the lowest
      // original index among exact matches
      if (mine && (tbest < best || tbest == best)) {
        int ti = 0x7fffffff;
        const bool multi = (twin & 256) != 0;
        const int s0 = multi ? 0 : (twin & 31), step = multi ? 1 : 32;
        for (int s = s0; s < navail; s += step) {
          const float4 c = sC[s];
          const float d = sqdist(c.x - q.x, c.y - q.y, c.z - q.z);
          if (__float_as_uint(d) == tbest) ti = min(ti, __float_as_int(c.w));
        }
        if (tbest < best || ti < bidx) {
          best = tbest;
          bidx = ti;
        }
      }
    }
  }

  // ---- finished?  Everything outside the halo of the query's block is at least `ext` cells away.
  bool done = false;
  if (active) {
    const float slx = 1e-4f + 1e-6f * (fabsf(p.ux) + (float)gx);  // rounding slack of the cell coordinates (grid.cuh)
    const float sly = 1e-4f + 1e-6f * (fabsf(p.uy) + (float)gy);
    const float slz = 1e-4f + 1e-6f * (fabsf(p.uz) + (float)gz);
    const float inf = __int_as_float(0x7f800000);
    float ext = inf;
    if (2 * bx - 1 > 0) ext = fminf(ext, p.ux - (float)(2 * bx - 1) - slx);
    if (2 * bx + 3 < gx) ext = fminf(ext, (float)(2 * bx + 3) - p.ux - slx);
    if (2 * by - 1 > 0) ext = fminf(ext, p.uy - (float)(2 * by - 1) - sly);
    if (2 * by + 3 < gy) ext = fminf(ext, (float)(2 * by + 3) - p.uy - sly);
    if (2 * bz - 1 > 0) ext = fminf(ext, p.uz - (float)(2 * bz - 1) - slz);
    if (2 * bz + 3 < gz) ext = fminf(ext, (float)(2 * bz + 3) - p.uz - slz);
    // in world units: a query far outside the grid is 1e20 CELLS away from it when the target cloud is tiny, and the
    // square of that overflows although the distance itself is ordinary
    const float ew = fmaxf(ext, 0.f) * h.s;
    done = bidx != 0x7fffffff && (ext == inf || __uint_as_float(best) < ew * ew * (1.f - 1e-5f));
  }
  if (present) {
    const int orig = __float_as_int(q.w);
    dist[orig] = __uint_as_float(best);  // final, or the bound the left-over pass starts from
    idx[orig] = bidx;
    const unsigned rest = __ballot_sync(__activemask(), !done);
    if (!done) {
      const int li = dir * b + cloud;
      const int leader = __ffs(rest) - 1;
      int pos = 0;
      if (lane == leader) pos = atomicAdd(W.count + li, __popc(rest));
      pos = __shfl_sync(rest, pos, leader) + __popc(rest & ((1u << lane) - 1u));
      ((dir ? W.list[1] : W.list[0]) + (size_t)cloud * nq)[pos] = orig;
    }
  }
}

// ------------------------------------------------------------------------------------------------ rest
// Completes the left-over queries, again with a warp as the unit: 32 consecutive entries of a (direction, cloud) list
// (appended by neighbouring lanes of the dense pass: neighbours in space), the provisional (distance, index) of the
// dense pass as each query's starting bound.
//   rows        the rows (y, z) of the target grid inside the rectangle the bounds of the 32 queries span are taken
//               32 at a time, LANES OVER ROWS: a lane tests its row against every query (broadcast from shared
//               memory) with the conservative lower bound of grid.cuh, in world units, and keeps the union of the
//               x-ranges the queries' bounds leave; a row no query needs contributes nothing.
//   candidates  the surviving ranges of a batch form one candidate list (prefix scan over the lanes), staged and
//               evaluated exactly like a halo of the dense pass (lanes over candidates, dq_tile), which tightens the
//               bounds for the next batch.
// A degenerate target grid is one cell holding the whole cloud: the same loop with a single row.  A non-finite query
// ends as (+inf, 0), the result of the reference's strict `<` scan on distances that are never smaller than +inf.
constexpr int kDrWarps = 4;
constexpr int kDrThreads = 32 * kDrWarps;
constexpr int kDrChunksPerList = 8;  // grid.x: CTAs past the end of a list leave at once

struct DrSmem {
  float4 cand[kDrWarps][kDqTile];
  float4 qry[kDrWarps][32];   // x, y, z, current bound (squared distance)
  float4 cell[kDrWarps][32];  // the query in cell units of the target grid
  int row_off[kDrWarps][32];
  int row_pos[kDrWarps][32];
};

__global__ void __launch_bounds__(kDrThreads)
chamfer_dense_rest_kernel(int b, int n, int m, DenseWs W, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                          float *__restrict__ dist1, float *__restrict__ dist2, int *__restrict__ idx1,
                          int *__restrict__ idx2) {
  __shared__ __align__(16) DrSmem S;
  const int li = blockIdx.y, dir = li >= b ? 1 : 0, cloud = dir ? li - b : li;
  const int cnt = __ldg(W.count + li);
  if ((int)(blockIdx.x * kDrThreads) >= cnt) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 *sC = S.cand[warp], *sQ = S.qry[warp], *sU = S.cell[warp];
  int *sOff = S.row_off[warp], *sPos = S.row_pos[warp];
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
      best = __float_as_uint(dist[orig]);
      bidx = idx[orig];
    }
    const CellPos p = cell_pos(h, q.x, q.y, q.z);
    const bool finite = fabsf(q.x) + fabsf(q.y) + fabsf(q.z) < 3.0e38f && fabsf(p.ux) + fabsf(p.uy) + fabsf(p.uz) < 3.0e38f;
    const bool active = present && finite;
    if (!(best <= 0x7f800000u)) best = 0x7f800000u, bidx = 0x7fffffff;  // a NaN bound is no bound
    const unsigned grp = __ballot_sync(0xffffffffu, active);
    __syncwarp();
    sQ[lane] = make_float4(q.x, q.y, q.z, active ? __uint_as_float(best) : -1.f);  // bound -1: needs no row
    sU[lane] = make_float4(p.ux, p.uy, p.uz, 0.f);
    __syncwarp();
    if (grp) {
      // rectangle of rows any query's bound reaches (the whole grid for an infinite bound); one row for a degenerate grid
      int ylo = 0, yhi = 0, zlo = 0, zhi = 0;
      if (h.valid) {
        float fy0 = (float)gy, fy1 = -1.f, fz0 = (float)gz, fz1 = -1.f;
        if (active) {
          const float r = sqrtf(__uint_as_float(best) * 1.0001f) * h.inv_s * 1.0001f + 1e-3f;  // cells (inf: everything)
          const float ry = r + 1e-4f + 1e-6f * (fabsf(p.uy) + (float)gy), rz = r + 1e-4f + 1e-6f * (fabsf(p.uz) + (float)gz);
          fy0 = fminf(fmaxf(floorf(p.uy - ry) - 1.f, 0.f), (float)(gy - 1));
          fy1 = fminf(fmaxf(floorf(p.uy + ry) + 1.f, 0.f), (float)(gy - 1));
          fz0 = fminf(fmaxf(floorf(p.uz - rz) - 1.f, 0.f), (float)(gz - 1));
          fz1 = fminf(fmaxf(floorf(p.uz + rz) + 1.f, 0.f), (float)(gz - 1));
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
        const int row = r0 + lane;
        if (row < nrows) {
          if (!h.valid) {
            a = 0, len = nt;
          } else {
            const int yy = ylo + row % ny, zz = zlo + row / ny;
            float fx0 = (float)gx, fx1 = -1.f;
            unsigned mq = grp;
            while (mq) {
              const int l = __ffs(mq) - 1;
              mq &= mq - 1;
              const float4 u = sU[l];
              const float bnd = sQ[l].w;
              const float gyy = cell_gap(u.y, yy, 1e-4f + 1e-6f * (fabsf(u.y) + (float)gy)) * h.s;
              const float gzz = cell_gap(u.z, zz, 1e-4f + 1e-6f * (fabsf(u.z) + (float)gz)) * h.s;
              const float lbyz = fmaf(gyy, gyy, gzz * gzz);
              if (!(lbyz * shr > bnd)) {
                // cells c of the row with (gap_x(c) s)^2 + lbyz <= bound / shr: a superset (the square root rounded
                // up, one extra cell at either end)
                const float xr = sqrtf(fmaxf(bnd * 1.0001f - lbyz * shr, 0.f)) * h.inv_s * 1.0001f + 1e-4f +
                                 1e-6f * (fabsf(u.x) + (float)gx) + 1e-3f;
                float f0 = fminf(fmaxf(floorf(u.x - xr) - 1.f, 0.f), (float)(gx - 1));
                float f1 = fminf(fmaxf(floorf(u.x + xr) + 1.f, 0.f), (float)(gx - 1));
                if (!(xr < inf)) f0 = 0.f, f1 = (float)(gx - 1);
                fx0 = fminf(fx0, f0), fx1 = fmaxf(fx1, f1);
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
        __syncwarp();
        // ---- lanes over candidates, one tile of the batch's list at a time
        for (int t0 = 0; t0 < total; t0 += kDqTile) {
          const int navail = min(total - t0, kDqTile);
          __syncwarp();
#pragma unroll 1
          for (int k0 = 0; k0 < kDqTile / 32 && 32 * k0 < navail; k0 += 4) {
            float4 cv[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const int sl = lane + 32 * (k0 + k);
              if (sl < navail) {
                const int g = t0 + sl;
                int r = sOff[16] <= g ? 16 : 0;
                r += sOff[r + 8] <= g ? 8 : 0;
                r += sOff[r + 4] <= g ? 4 : 0;
                r += sOff[r + 2] <= g ? 2 : 0;
                r += sOff[r + 1] <= g ? 1 : 0;
                cv[k] = __ldg(T + sPos[r] + (g - sOff[r]));
              }
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
              if (lane + 32 * (k0 + k) < navail) sC[lane + 32 * (k0 + k)] = cv[k];
          }
          __syncwarp();
          uint32_t tbest = 0x7fffffffu;
          int twin = 0;
          const int k2 = (navail + 63) >> 6;
          if (k2 == 1) dq_tile<1>(sC, sQ, grp, lane, navail, tbest, twin);
          else if (k2 == 2) dq_tile<2>(sC, sQ, grp, lane, navail, tbest, twin);
          else if (k2 == 3) dq_tile<3>(sC, sQ, grp, lane, navail, tbest, twin);
          else dq_tile<4>(sC, sQ, grp, lane, navail, tbest, twin);
          if (active && tbest <= best) {
            int ti = 0x7fffffff;
            const bool multi = (twin & 256) != 0;
            const int s0 = multi ? 0 : (twin & 31), step = multi ? 1 : 32;
            for (int s = s0; s < navail; s += step) {
              const float4 c = sC[s];
              const float d = sqdist(c.x - q.x, c.y - q.y, c.z - q.z);
              if (__float_as_uint(d) == tbest) ti = min(ti, __float_as_int(c.w));
            }
            if (tbest < best || ti < bidx) {
              best = tbest;
              bidx = ti;
            }
          }
        }
        __syncwarp();
        if (active) sQ[lane].w = __uint_as_float(best);  // the tightened bound prunes the following batches
        __syncwarp();
      }
    }
    if (present) {
      dist[orig] = __uint_as_float(best);
      idx[orig] = bidx == 0x7fffffff ? 0 : bidx;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host
bool chamfer_dense_supported(int b, int n, int m) {
  return b > 0 && b <= 32767 && n >= kDenseMinPts && m >= kDenseMinPts && n <= kDenseMaxPts && m <= kDenseMaxPts;
}

size_t chamfer_dense_workspace_bytes(int b, int n, int m) { return dense_plan(b, n, m, nullptr, nullptr); }

template <int PARTS>
static int dense_build_launch(int b, int n, int m, const float *xyz1, const float *xyz2, const DenseWs &W,
                              cudaStream_t s) {
  const size_t smem = sizeof(int) * (size_t)(kDbPerO + kDbPerB + 2 * kDbPts) * kDbThreads;
  static size_t granted[kMaxDevices];
  const int rc = grant_dyn_smem(chamfer_dense_build_kernel<PARTS>, smem, granted, 0);
  if (rc) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(b * 2 * PARTS));
  cfg.blockDim = dim3(kDbThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * PARTS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, chamfer_dense_build_kernel<PARTS>, b, n, m, xyz1, xyz2, W);
  return e == cudaSuccess ? MVP_OK : (int)e;
}

int chamfer_dense_launch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1, float *dist2,
                         int *idx1, int *idx2, void *ws, size_t ws_bytes, cudaStream_t s) {
  DenseWs W;
  if (ws_bytes < dense_plan(b, n, m, ws, &W)) return MVP_ERR_WORKSPACE;
  int rc = std::max(n, m) >= kDbSplitPts ? dense_build_launch<2>(b, n, m, xyz1, xyz2, W, s)
                                         : dense_build_launch<1>(b, n, m, xyz1, xyz2, W, s);
  if (rc) return rc;
  const long long warps = (long long)b * (((n + 31) >> 5) + ((m + 31) >> 5));
  chamfer_dense_query_kernel<<<(unsigned)((warps + kDqWarps - 1) / kDqWarps), kDqThreads, 0, s>>>(b, n, m, W, dist1, dist2,
                                                                                                idx1, idx2);
  chamfer_dense_rest_kernel<<<dim3(kDrChunksPerList, 2 * b), kDrThreads, 0, s>>>(b, n, m, W, xyz1, xyz2, dist1, dist2, idx1,
                                                                              idx2);
  count_launch(3);
  return launch_status();
}

}  // namespace mvp
