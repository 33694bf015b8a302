// Earth mover's distance (parallel auction) for sm_100a.
//
// Replaces the reference's host loop of 7 kernel launches per auction round
// (utils/metrics/EMD/emd_cuda.cu:256-269; kernels clear/calc_unass_cnt/calc_unass_cnt_sum/calc_unass_idx/
// Bid/GetMax/Assign :23-215, CalcDist :217-226) with ONE persistent launch: a thread-block CLUSTER owns a
// cloud for the whole auction, the targets and their prices stay resident in shared memory as float4
// (x, y, z, price), and the rounds are separated by cluster barriers instead of kernel boundaries.
// An auction that has converged (no unassigned source left) leaves the loop early — later rounds are
// no-ops in the reference too.
//
// Arithmetic kept bit-for-bit:
//   value  = (float)(3.0 - (double)sqrtf(s) - (double)price),  s = fma(dz,dz,fma(dx,dx,dy*dy))   (:142-146)
//   top-2 with strict `>` and duplicates counted                                               (:147-154,165-173)
//   best_i among equal values = first in the reference's (cooperating thread, tile, k) scan order, which
//          depends on thread_per_unass = 1024 / ceil(unassigned / (n/1024))                    (:108-110,136-139)
//   increment = best - better + eps; max per target; a bidder within +-1e-6 (double) of the max may
//          win — the reference lets the last writer win (:188-191); here the HIGHEST source index wins
//   winner evicts, price += its own increment, max_increments := -1e9; last round assigns everyone  (:196-215)
#include <cooperative_groups.h>

#include "common.cuh"
#include "grid.cuh"

namespace cg = cooperative_groups;

namespace mvp {

constexpr int kEmdThreads = 1024;
constexpr int kEmdMaxTile = 8192;  // targets resident per tile: 8192 * 16 B = 128 KB

struct EmdState {  // per-cloud slices of the caller's workspace
  unsigned long long *max_idx;  // (round+1) << 32 | source
  float *price;
  int *assignment_inv;
  int *bid;
  float *bid_inc;
  float *max_inc;
};

struct BidTop {
  float best, better;
  int bi;
};

struct KeyParams {
  int n, tpu;  // tpu = the reference's thread_per_unass for this cloud and round
};

// Position of target k in the reference's scan order: (segment inside its 2048-tile, k).
__device__ __forceinline__ int ref_segment(int k, const KeyParams &kp) {
  const int k2 = k & ~2047;
  const int end_k = min(2048, kp.n - k2);
  const int delta = (end_k + kp.tpu - 1) / kp.tpu;
  return (k - k2) / delta;
}
__device__ __forceinline__ bool scans_before(int a, int b, const KeyParams &kp) {
  if (b < 0) return true;
  if (a < 0) return false;
  const int sa = ref_segment(a, kp), sb = ref_segment(b, kp);
  return sa < sb || (sa == sb && a < b);
}

__device__ __forceinline__ void merge_top(BidTop &s, const BidTop &o, const KeyParams &kp) {
  if (o.best > s.best) {
    s.better = fmaxf(s.best, o.better);
    s.best = o.best;
    s.bi = o.bi;
  } else if (o.best == s.best) {
    s.better = s.best;
    if (scans_before(o.bi, s.bi, kp)) s.bi = o.bi;
  } else {
    s.better = fmaxf(s.better, o.best);
  }
}

__device__ __forceinline__ void atomic_max_float(float *addr, float val) {
  // emd_cuda.cu:10-21.  Non-negative values order like their bit patterns as signed ints, and any of them
  // beats the negative reset value -1e9; only a negative `val` (eps < 0) needs the CAS loop.
  if (val >= 0.f) {
    atomicMax(reinterpret_cast<int *>(addr), __float_as_int(val));
  } else {
    int ret = __float_as_int(*reinterpret_cast<volatile float *>(addr));
    while (val > __int_as_float(ret)) {
      const int old = ret;
      if ((ret = atomicCAS(reinterpret_cast<int *>(addr), old, __float_as_int(val))) == old) break;
    }
  }
}

__global__ void __launch_bounds__(kEmdThreads, 1)
emd_auction_kernel(int n, int tile_cap, float eps, int iters, const float *__restrict__ xyz1_all,
                   const float *__restrict__ xyz2_all, float *__restrict__ dist_all,
                   int *__restrict__ assignment_all, unsigned char *__restrict__ ws_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int cloud = blockIdx.x / C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  float4 *tgt = reinterpret_cast<float4 *>(smem_raw);                       // tile_cap entries
  int *list = reinterpret_cast<int *>(smem_raw + (size_t)tile_cap * 16);    // n / C entries
  __shared__ int s_cnt;
  __shared__ int s_total;  // read by the other CTAs of the cluster through DSMEM
  __shared__ BidTop s_merge[32];

  const float *xyz1 = xyz1_all + (size_t)cloud * n * 3;
  const float *xyz2 = xyz2_all + (size_t)cloud * n * 3;
  float *dist = dist_all + (size_t)cloud * n;
  int *assignment = assignment_all + (size_t)cloud * n;
  EmdState st;
  {
    unsigned char *w = ws_all + (size_t)cloud * n * 28;
    st.max_idx = reinterpret_cast<unsigned long long *>(w);
    st.price = reinterpret_cast<float *>(w + (size_t)n * 8);
    st.assignment_inv = reinterpret_cast<int *>(w + (size_t)n * 12);
    st.bid = reinterpret_cast<int *>(w + (size_t)n * 16);
    st.bid_inc = reinterpret_cast<float *>(w + (size_t)n * 20);
    st.max_inc = reinterpret_cast<float *>(w + (size_t)n * 24);
  }
  const int per = n / C, lo = rank * per, hi = lo + per;

  // State initialisation of emd_module.py:54-65 (own slice).
  for (int j = lo + tid; j < hi; j += kEmdThreads) {
    assignment[j] = -1;
    st.assignment_inv[j] = -1;
    st.price[j] = 0.f;
    st.max_inc[j] = 0.f;
    st.max_idx[j] = 0ull;
  }
  const int ntiles = (n + tile_cap - 1) / tile_cap;
  if (ntiles == 1)
    for (int k = tid; k < n; k += kEmdThreads)
      tgt[k] = make_float4(__ldg(xyz2 + k * 3 + 0), __ldg(xyz2 + k * 3 + 1), __ldg(xyz2 + k * 3 + 2), 0.f);
  cluster.sync();

  for (int it = 0; it < iters; it++) {
    const bool last = (it == iters - 1);
    // ---- list the unassigned sources of the own slice (order is irrelevant, as in calc_unass_idx :85-93)
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int j0 = lo; j0 < hi; j0 += kEmdThreads) {
      const int j = j0 + tid;
      const bool un = (j < hi) && (__ldcg(assignment + j) == -1);
      const unsigned mask = __ballot_sync(0xffffffffu, un);
      if (mask) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_cnt, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (un) list[base + __popc(mask & ((1u << lane) - 1u))] = j;
      }
    }
    __syncthreads();
    const int ucnt = s_cnt;
    if (tid == 0) s_total = ucnt;
    cluster.sync();
    int total = 0;
    for (int r = 0; r < C; r++) total += *cluster.map_shared_rank(&s_total, r);
    if (total == 0) break;  // converged: every later round is a no-op
    KeyParams kp;
    kp.n = n;
    {
      const int block_cnt = n / 1024;
      const int unass_per_block = (total + block_cnt - 1) / block_cnt;
      kp.tpu = 1024 / unass_per_block;
    }

    // ---- Bid
    int tpp = 1;  // threads per source: largest power of two <= kEmdThreads / ucnt
    while (ucnt > 0 && tpp * 2 * ucnt <= kEmdThreads) tpp *= 2;
    const int ppp = kEmdThreads / tpp;  // sources per pass
    const int passes = (ucnt + ppp - 1) / ppp;
    if (ntiles == 1) {
      for (int k = tid; k < n; k += kEmdThreads) tgt[k].w = __ldcg(st.price + k);
      __syncthreads();
    }
    for (int pass = 0; pass < passes; pass++) {
      const int pi = pass * ppp + tid / tpp;
      const int g = tid % tpp;
      const bool valid = pi < ucnt;
      int src = -1;
      float x1 = 0, y1 = 0, z1 = 0;
      if (valid) {
        src = list[pi];
        x1 = __ldg(xyz1 + src * 3 + 0);
        y1 = __ldg(xyz1 + src * 3 + 1);
        z1 = __ldg(xyz1 + src * 3 + 2);
      }
      BidTop top;
      top.best = -1e9f;
      top.better = -1e9f;
      top.bi = -1;
      for (int t = 0; t < ntiles; t++) {
        const int k2 = t * tile_cap;
        const int tcnt = min(tile_cap, n - k2);
        if (ntiles > 1) {
          __syncthreads();
          for (int k = tid; k < tcnt; k += kEmdThreads)
            tgt[k] = make_float4(__ldg(xyz2 + (k2 + k) * 3 + 0), __ldg(xyz2 + (k2 + k) * 3 + 1),
                                 __ldg(xyz2 + (k2 + k) * 3 + 2), __ldcg(st.price + k2 + k));
          __syncthreads();
        }
        if (valid) {
          for (int k = g; k < tcnt; k += tpp) {
            const float4 q = tgt[k];
            const float s = sqdist(q.x - x1, q.y - y1, q.z - z1);
            const float d = (float)(3.0 - (double)__fsqrt_rn(s) - (double)q.w);
            if (d > top.best) {
              top.better = top.best;
              top.best = d;
              top.bi = k2 + k;
            } else if (d == top.best) {
              top.better = d;
              if (scans_before(k2 + k, top.bi, kp)) top.bi = k2 + k;
            } else if (d > top.better) {
              top.better = d;
            }
          }
        }
      }
      // merge the tpp partial results of each source
      const int wspan = tpp < 32 ? tpp : 32;
      for (int off = 1; off < wspan; off <<= 1) {
        BidTop o;
        o.best = __shfl_xor_sync(0xffffffffu, top.best, off);
        o.better = __shfl_xor_sync(0xffffffffu, top.better, off);
        o.bi = __shfl_xor_sync(0xffffffffu, top.bi, off);
        merge_top(top, o, kp);
      }
      if (tpp > 32) {  // one source spans tpp/32 whole warps
        __syncthreads();
        if (lane == 0) s_merge[warp] = top;
        __syncthreads();
        const int wpp = tpp / 32;
        if (lane == 0 && (warp % wpp) == 0) {
          for (int w = warp + 1; w < warp + wpp; w++) merge_top(top, s_merge[w], kp);
        }
      }
      if (valid && g == 0) {
        const float inc = top.best - top.better + eps;
        st.bid[src] = top.bi;
        st.bid_inc[src] = inc;
        atomic_max_float(st.max_inc + top.bi, inc);
      }
    }
    cluster.sync();

    // ---- GetMax: bidders within +-1e-6 of the target's maximum; highest source index wins
    const unsigned long long round_tag = (unsigned long long)(it + 1) << 32;
    for (int u = tid; u < ucnt; u += kEmdThreads) {
      const int j = list[u];
      const int bid_id = __ldcg(st.bid + j);
      const float bid_inc = __ldcg(st.bid_inc + j);
      const float max_inc = __ldcg(st.max_inc + bid_id);
      if ((double)bid_inc - 1e-6 <= (double)max_inc && (double)max_inc <= (double)bid_inc + 1e-6)
        atomicMax(st.max_idx + bid_id, round_tag | (unsigned)j);
    }
    cluster.sync();

    // ---- Assign
    for (int u = tid; u < ucnt; u += kEmdThreads) {
      const int j = list[u];
      const int bid_id = __ldcg(st.bid + j);
      if (last) {
        assignment[j] = bid_id;  // everyone is assigned, nobody evicted (:201-207); price is no output
      } else if (__ldcg(st.max_idx + bid_id) == (round_tag | (unsigned)j)) {
        const float bid_inc = __ldcg(st.bid_inc + j);
        const int ass_inv = __ldcg(st.assignment_inv + bid_id);
        if (ass_inv != -1) assignment[ass_inv] = -1;
        st.assignment_inv[bid_id] = j;
        assignment[j] = bid_id;
        st.price[bid_id] = __ldcg(st.price + bid_id) + bid_inc;
        st.max_inc[bid_id] = -1e9f;
      }
    }
    cluster.sync();
  }

  // ---- CalcDist (:217-226), own slice.  assignment == -1 can only remain when iters == 0.
  for (int j = lo + tid; j < hi; j += kEmdThreads) {
    const int k = __ldcg(assignment + j);
    float d = 0.f;
    if (k >= 0)
      d = sqdist(__ldg(xyz1 + j * 3 + 0) - __ldg(xyz2 + k * 3 + 0), __ldg(xyz1 + j * 3 + 1) - __ldg(xyz2 + k * 3 + 1),
                 __ldg(xyz1 + j * 3 + 2) - __ldg(xyz2 + k * 3 + 2));
    dist[j] = d;
  }
  cluster.sync();  // no CTA may exit while a peer can still read its shared memory (s_total)
}

// ------------------------------------------------------------------------------------------------------------
// Grid-pruned auction (n <= 8192: the whole target cloud, its prices and a uniform grid over it fit in shared
// memory).  Same rounds, same arithmetic and the same tie rules as emd_auction_kernel above — only the Bid search
// changes: instead of scanning all n targets, a source visits cubes of grid cells of growing radius around itself
// and stops when no unvisited target can reach its current second-best value:
//     value_j = 3 - |x_i - y_j| - price_j  <=  3 - dist(x_i, cell) - min_price,
// so a row of cells (or everything outside the visited cube) is skipped when that bound, taken conservatively
// (1e-5 relative + 1e-5 absolute margins against the few-ulp error of the float/double evaluation), is below
// `better`.  Skipped targets have value < better <= best, hence cannot change best, better or best_i: the bids are
// bit-identical to the full scan.  Late in an auction prices are a few eps and a bid needs ~10^2 targets, not n.
#ifndef MVP_EMD_GRID_PPC
#define MVP_EMD_GRID_PPC 4
#endif
constexpr int kEmdGridPPC = MVP_EMD_GRID_PPC;  // target points per grid cell
constexpr int kEmdGridMaxN = 8192;
#ifndef MVP_EMD_FULLSCAN_EVALS
#define MVP_EMD_FULLSCAN_EVALS 48
#endif
constexpr int kEmdFullScanEvals = MVP_EMD_FULLSCAN_EVALS;  // per-thread evaluations up to which a full scan is used

__global__ void __launch_bounds__(kEmdThreads, 1)
emd_auction_grid_kernel(int n, int cap, float eps, int iters, const float *__restrict__ xyz1_all,
                        const float *__restrict__ xyz2_all, float *__restrict__ dist_all,
                        int *__restrict__ assignment_all, unsigned char *__restrict__ ws_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int cloud = blockIdx.x / C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float inf = __int_as_float(0x7f800000);

  float4 *tgt = reinterpret_cast<float4 *>(smem_raw);                    // n: (x, y, z, price), sorted by cell
  int *orig = reinterpret_cast<int *>(smem_raw + (size_t)n * 16);        // n: original index of a sorted target
  int *T = orig + n;                                                     // cap + 1 (padded to 4): cell starts
  int *list = T + ((cap + 1 + 3) & ~3);                                  // n / C: unassigned sources of this CTA
  __shared__ int s_cnt;
  __shared__ int s_total;  // read by the other CTAs of the cluster through DSMEM
  __shared__ float s_red[6][32];
  __shared__ int s_int[32];
  __shared__ GridHdr s_hdr;
  __shared__ float s_pmin;
  __shared__ BidTop s_merge[32];

  const float *xyz1 = xyz1_all + (size_t)cloud * n * 3;
  const float *xyz2 = xyz2_all + (size_t)cloud * n * 3;
  float *dist = dist_all + (size_t)cloud * n;
  int *assignment = assignment_all + (size_t)cloud * n;
  EmdState st;
  {
    unsigned char *w = ws_all + (size_t)cloud * n * 28;
    st.max_idx = reinterpret_cast<unsigned long long *>(w);
    st.price = reinterpret_cast<float *>(w + (size_t)n * 8);
    st.assignment_inv = reinterpret_cast<int *>(w + (size_t)n * 12);
    st.bid = reinterpret_cast<int *>(w + (size_t)n * 16);
    st.bid_inc = reinterpret_cast<float *>(w + (size_t)n * 20);
    st.max_inc = reinterpret_cast<float *>(w + (size_t)n * 24);
  }
  const int per = n / C, lo_j = rank * per, hi_j = lo_j + per;

  // State initialisation of emd_module.py:54-65 (own slice).
  for (int j = lo_j + tid; j < hi_j; j += kEmdThreads) {
    assignment[j] = -1;
    st.assignment_inv[j] = -1;
    st.price[j] = 0.f;
    st.max_inc[j] = 0.f;
    st.max_idx[j] = 0ull;
  }

  // ---- grid over the targets (every CTA of the cluster builds its own copy; the order inside a cell differs
  // between copies, which is why global state is indexed by ORIGINAL target index)
  {
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    int fin = 1;
    for (int k = tid; k < n; k += kEmdThreads) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const float v = __ldg(xyz2 + (size_t)k * 3 + a);
        lo[a] = fminf(lo[a], v);
        hi[a] = fmaxf(hi[a], v);
        fin &= (fabsf(v) <= 3.0e38f) ? 1 : 0;
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
      s_int[warp] = fin;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kEmdThreads / 32; w++) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
          lo[a] = fminf(lo[a], s_red[a][w]);
          hi[a] = fmaxf(hi[a], s_red[3 + a][w]);
        }
        fin &= s_int[w];
      }
      s_hdr = grid_header(lo, hi, fin, cap);
    }
    __syncthreads();
  }
  const GridHdr h = s_hdr;
  const int gx = h.g[0], gy = h.g[1], gz = h.g[2], ncell = h.ncell;
  auto cell_of = [&](float x, float y, float z) {
    if (!h.valid) return 0;
    const int cx = cell_coord((x - h.lo[0]) * h.inv_s, gx);
    const int cy = cell_coord((y - h.lo[1]) * h.inv_s, gy);
    const int cz = cell_coord((z - h.lo[2]) * h.inv_s, gz);
    return (cz * gy + cy) * gx + cx;
  };
  for (int c = tid; c <= ncell; c += kEmdThreads) T[c] = 0;
  __syncthreads();
  for (int k = tid; k < n; k += kEmdThreads)
    atomicAdd(&T[cell_of(__ldg(xyz2 + k * 3 + 0), __ldg(xyz2 + k * 3 + 1), __ldg(xyz2 + k * 3 + 2)) + 1], 1);
  __syncthreads();
  {  // exclusive scan: T[c + 1] := start of cell c (it then serves as the scatter cursor and ends as start[c + 1])
    const int chunk = (ncell + kEmdThreads - 1) / kEmdThreads;
    const int c0 = min(tid * chunk, ncell), c1 = min(c0 + chunk, ncell);
    int sum = 0;
    for (int c = c0; c < c1; c++) sum += T[c + 1];
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    if (lane == 31) s_int[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = s_int[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= off) v += u;
      }
      s_int[lane] = v;
    }
    __syncthreads();
    int run = incl - sum + (warp ? s_int[warp - 1] : 0);
    for (int c = c0; c < c1; c++) {
      const int cnt = T[c + 1];
      T[c + 1] = run;
      run += cnt;
    }
  }
  __syncthreads();
  for (int k = tid; k < n; k += kEmdThreads) {
    const float x = __ldg(xyz2 + k * 3 + 0), y = __ldg(xyz2 + k * 3 + 1), z = __ldg(xyz2 + k * 3 + 2);
    const int pos = atomicAdd(&T[cell_of(x, y, z) + 1], 1);
    tgt[pos] = make_float4(x, y, z, 0.f);
    orig[pos] = k;
  }
  cluster.sync();  // T[c] = first sorted position of cell c, T[ncell] = n; state initialised cluster-wide

  const float s2 = h.s * h.s * (1.f - 1e-5f);  // cells^2 -> squared distance, rounded down generously

  // current prices (global, indexed by original target) into the sorted copy; their minimum bounds every unvisited
  // target's value
  auto refresh_prices = [&]() {
  {
    float pm = inf;
    for (int pos = tid; pos < n; pos += kEmdThreads) {
      const float p = __ldcg(st.price + orig[pos]);
      tgt[pos].w = p;
      pm = fminf(pm, p);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) pm = fminf(pm, __shfl_xor_sync(0xffffffffu, pm, off));
    if (lane == 0) s_red[0][warp] = pm;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kEmdThreads / 32; w++) pm = fminf(pm, s_red[0][w]);
      s_pmin = pm;
    }
    __syncthreads();
  }
  };
  auto bid_phase = [&](const int ucnt, const KeyParams kp) {
  const float vmax = 3.0f - s_pmin;  // no target's value exceeds this (up to rounding, covered by the margins)
  // squared distance beyond which a target cannot reach `better`; negative = nothing can
  auto reach2 = [&](float better) {
    const float t = (vmax - better) * (1.f + 1e-5f) + 1e-5f;
    return t < 0.f ? -1.f : t * t;
  };

  // ---- Bid
  // Few unassigned sources (the long tail of an auction): spreading a full scan of the shared-memory targets over
  // all 1024 threads costs n * ucnt / 1024 evaluations per thread and no search overhead — cheaper than a grid
  // search by one warp per source once that is a few dozen.
  if ((long long)ucnt * n <= (long long)kEmdThreads * kEmdFullScanEvals) {
    int tpp = 1;  // threads per source: largest power of two <= kEmdThreads / ucnt
    while (ucnt > 0 && tpp * 2 * ucnt <= kEmdThreads) tpp *= 2;
    const int pi = tid / tpp, g = tid % tpp;
    const bool valid = pi < ucnt;
    int src = -1;
    float x1 = 0, y1 = 0, z1 = 0;
    if (valid) {
      src = list[pi];
      x1 = __ldg(xyz1 + src * 3 + 0);
      y1 = __ldg(xyz1 + src * 3 + 1);
      z1 = __ldg(xyz1 + src * 3 + 2);
    }
    BidTop top;
    top.best = -1e9f;
    top.better = -1e9f;
    top.bi = -1;
    if (valid) {
      for (int pos = g; pos < n; pos += tpp) {
        const float4 q = tgt[pos];
        const float s = sqdist(q.x - x1, q.y - y1, q.z - z1);
        const float d = (float)(3.0 - (double)__fsqrt_rn(s) - (double)q.w);
        if (d > top.best) {
          top.better = top.best;
          top.best = d;
          top.bi = orig[pos];
        } else if (d == top.best) {
          top.better = d;
          const int k = orig[pos];
          if (scans_before(k, top.bi, kp)) top.bi = k;
        } else if (d > top.better) {
          top.better = d;
        }
      }
    }
    const int wspan = tpp < 32 ? tpp : 32;
    for (int off = 1; off < wspan; off <<= 1) {
      BidTop o;
      o.best = __shfl_xor_sync(0xffffffffu, top.best, off);
      o.better = __shfl_xor_sync(0xffffffffu, top.better, off);
      o.bi = __shfl_xor_sync(0xffffffffu, top.bi, off);
      merge_top(top, o, kp);
    }
    if (tpp > 32) {  // one source spans tpp/32 whole warps
      if (lane == 0) s_merge[warp] = top;
      __syncthreads();
      const int wpp = tpp / 32;
      if (lane == 0 && (warp % wpp) == 0)
        for (int w = warp + 1; w < warp + wpp; w++) merge_top(top, s_merge[w], kp);
    }
    if (valid && g == 0) {
      const float inc = top.best - top.better + eps;
      st.bid[src] = top.bi;
      st.bid_inc[src] = inc;
      atomic_max_float(st.max_inc + top.bi, inc);
    }
  } else {
  int tpp = 1;  // lanes per source: largest power of two <= min(32, kEmdThreads / ucnt)
  while (ucnt > 0 && tpp < 32 && tpp * 2 * ucnt <= kEmdThreads) tpp *= 2;
  const int ppp = kEmdThreads / tpp;  // sources per pass
  const int passes = (ucnt + ppp - 1) / ppp;
  for (int pass = 0; pass < passes; pass++) {
    const int pi = pass * ppp + tid / tpp;
    const int g = tid % tpp;
    const bool valid = pi < ucnt;
    int src = -1;
    float x1 = 0, y1 = 0, z1 = 0;
    if (valid) {
      src = list[pi];
      x1 = __ldg(xyz1 + src * 3 + 0);
      y1 = __ldg(xyz1 + src * 3 + 1);
      z1 = __ldg(xyz1 + src * 3 + 2);
    }
    BidTop top;
    top.best = -1e9f;
    top.better = -1e9f;
    top.bi = -1;
    const float ux = (x1 - h.lo[0]) * h.inv_s, uy = (y1 - h.lo[1]) * h.inv_s, uz = (z1 - h.lo[2]) * h.inv_s;
    const int cx = cell_coord(ux, gx), cy = cell_coord(uy, gy), cz = cell_coord(uz, gz);
    const float slx = 1e-4f + 1e-6f * (fabsf(ux) + (float)gx);
    const float sly = 1e-4f + 1e-6f * (fabsf(uy) + (float)gy);
    const float slz = 1e-4f + 1e-6f * (fabsf(uz) + (float)gz);
    // a source whose cell coordinates are not finite cannot prune: it scans the whole grid ring by ring
    const bool prunable = h.valid && (fabsf(ux) + fabsf(uy) + fabsf(uz) < 1e30f);

    // Lane g of a source's group owns the sorted positions == g (mod tpp), whatever range it is handed: the lanes
    // clip their ranges independently (each with its own, conservative, `better`), so the split must not depend
    // on where a lane's range starts.
    auto scan = [&](int a, int e) {
      for (int pos = a + ((g - a) & (tpp - 1)); pos < e; pos += tpp) {
        const float4 q = tgt[pos];
        const float s = sqdist(q.x - x1, q.y - y1, q.z - z1);
        const float d = (float)(3.0 - (double)__fsqrt_rn(s) - (double)q.w);
        if (d > top.best) {
          top.better = top.best;
          top.best = d;
          top.bi = orig[pos];
        } else if (d == top.best) {
          top.better = d;
          const int k = orig[pos];
          if (scans_before(k, top.bi, kp)) top.bi = k;
        } else if (d > top.better) {
          top.better = d;
        }
      }
    };

    BidTop merged = top;
    float known = -1e9f;   // a lower bound of the group's `better` (merged at the end of the previous ring)
    bool finished = !valid;
    int r = 0;
    while (__any_sync(0xffffffffu, !finished)) {
      if (!finished) {
        for (int dz = -r; dz <= r; dz++) {
          const int zz = cz + dz;
          if (zz < 0 || zz >= gz) continue;
          const float gzz = prunable ? cell_gap(uz, zz, slz) : 0.f;
          for (int dy = -r; dy <= r; dy++) {
            const int yy = cy + dy;
            if (yy < 0 || yy >= gy) continue;
            const float gyy = prunable ? cell_gap(uy, yy, sly) : 0.f;
            const float lbyz = fmaf(gyy, gyy, gzz * gzz);
#ifdef MVP_EMD_DEBUG_NOPRUNE
            const float lim = 3e38f;
#else
            const float lim = reach2(fmaxf(known, top.better));
#endif
            if (lbyz * s2 > lim) continue;
            const int base = (zz * gy + yy) * gx;
            const bool shell = (dz == -r || dz == r || dy == -r || dy == r);
            if (shell) {
              int x0 = max(cx - r, 0), x1c = min(cx + r, gx - 1);
              if (prunable) {
                while (x0 <= x1c) {
                  const float gg = cell_gap(ux, x0, slx);
                  if (fmaf(gg, gg, lbyz) * s2 > lim) x0++; else break;
                }
                while (x1c >= x0) {
                  const float gg = cell_gap(ux, x1c, slx);
                  if (fmaf(gg, gg, lbyz) * s2 > lim) x1c--; else break;
                }
              }
              if (x0 <= x1c) scan(T[base + x0], T[base + x1c + 1]);
            } else {  // interior row: only its two new end cells
              if (cx - r >= 0) {
                const float gg = prunable ? cell_gap(ux, cx - r, slx) : 0.f;
                if (!(fmaf(gg, gg, lbyz) * s2 > lim)) scan(T[base + cx - r], T[base + cx - r + 1]);
              }
              if (cx + r < gx) {
                const float gg = prunable ? cell_gap(ux, cx + r, slx) : 0.f;
                if (!(fmaf(gg, gg, lbyz) * s2 > lim)) scan(T[base + cx + r], T[base + cx + r + 1]);
              }
            }
          }
        }
      }
      // merge the partial results of the tpp lanes of each source (all lanes of the warp take part)
      merged = top;
      for (int off = 1; off < tpp; off <<= 1) {
        BidTop o;
        o.best = __shfl_xor_sync(0xffffffffu, merged.best, off);
        o.better = __shfl_xor_sync(0xffffffffu, merged.better, off);
        o.bi = __shfl_xor_sync(0xffffffffu, merged.bi, off);
        merge_top(merged, o, kp);
      }
      if (!finished) {
        known = merged.better;
        // everything outside the cube of radius r is at least `ext` cells away
        float ext = inf;
        if (cx - r > 0) ext = fminf(ext, ux - (float)(cx - r) - slx);
        if (cx + r + 1 < gx) ext = fminf(ext, (float)(cx + r + 1) - ux - slx);
        if (cy - r > 0) ext = fminf(ext, uy - (float)(cy - r) - sly);
        if (cy + r + 1 < gy) ext = fminf(ext, (float)(cy + r + 1) - uy - sly);
        if (cz - r > 0) ext = fminf(ext, uz - (float)(cz - r) - slz);
        if (cz + r + 1 < gz) ext = fminf(ext, (float)(cz + r + 1) - uz - slz);
        const bool covered = (cx - r <= 0) && (cx + r + 1 >= gx) && (cy - r <= 0) && (cy + r + 1 >= gy) &&
                             (cz - r <= 0) && (cz + r + 1 >= gz);
        ext = fmaxf(ext, 0.f);
#ifdef MVP_EMD_DEBUG_NOTERM
        if (covered) finished = true;
#else
        if (covered || (prunable && ext * ext * s2 > reach2(known))) finished = true;
#endif
        r++;
      }
    }
    if (valid && g == 0) {
      const float inc = merged.best - merged.better + eps;
      st.bid[src] = merged.bi;
      st.bid_inc[src] = inc;
      atomic_max_float(st.max_inc + merged.bi, inc);
    }
  }
  }  // grid search
  };

  for (int it = 0; it < iters; it++) {
    const bool last = (it == iters - 1);
    // ---- list the unassigned sources of the own slice (order is irrelevant, as in calc_unass_idx :85-93)
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int j0 = lo_j; j0 < hi_j; j0 += kEmdThreads) {
      const int j = j0 + tid;
      const bool un = (j < hi_j) && (__ldcg(assignment + j) == -1);
      const unsigned mask = __ballot_sync(0xffffffffu, un);
      if (mask) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_cnt, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (un) list[base + __popc(mask & ((1u << lane) - 1u))] = j;
      }
    }
    __syncthreads();
    const int ucnt = s_cnt;
    if (tid == 0) s_total = ucnt;
    cluster.sync();
    int total = 0;
    for (int r = 0; r < C; r++) total += *cluster.map_shared_rank(&s_total, r);
    if (total == 0) break;  // converged: every later round is a no-op
    KeyParams kp;
    kp.n = n;
    {
      const int block_cnt = n / 1024;
      const int unass_per_block = (total + block_cnt - 1) / block_cnt;
      kp.tpu = 1024 / unass_per_block;
    }

    refresh_prices();
    bid_phase(ucnt, kp);
    cluster.sync();

    // ---- GetMax: bidders within +-1e-6 of the target's maximum; highest source index wins
    const unsigned long long round_tag = (unsigned long long)(it + 1) << 32;
    for (int u = tid; u < ucnt; u += kEmdThreads) {
      const int j = list[u];
      const int bid_id = __ldcg(st.bid + j);
      const float bid_inc = __ldcg(st.bid_inc + j);
      const float max_inc = __ldcg(st.max_inc + bid_id);
      if ((double)bid_inc - 1e-6 <= (double)max_inc && (double)max_inc <= (double)bid_inc + 1e-6)
        atomicMax(st.max_idx + bid_id, round_tag | (unsigned)j);
    }
    cluster.sync();

    // ---- Assign
    for (int u = tid; u < ucnt; u += kEmdThreads) {
      const int j = list[u];
      const int bid_id = __ldcg(st.bid + j);
      if (last) {
        assignment[j] = bid_id;  // everyone is assigned, nobody evicted (:201-207); price is no output
      } else if (__ldcg(st.max_idx + bid_id) == (round_tag | (unsigned)j)) {
        const float bid_inc = __ldcg(st.bid_inc + j);
        const int ass_inv = __ldcg(st.assignment_inv + bid_id);
        if (ass_inv != -1) assignment[ass_inv] = -1;
        st.assignment_inv[bid_id] = j;
        assignment[j] = bid_id;
        st.price[bid_id] = __ldcg(st.price + bid_id) + bid_inc;
        st.max_inc[bid_id] = -1e9f;
      }
    }
    cluster.sync();
  }

  // ---- CalcDist (:217-226), own slice.  assignment == -1 can only remain when iters == 0.
  for (int j = lo_j + tid; j < hi_j; j += kEmdThreads) {
    const int k = __ldcg(assignment + j);
    float d = 0.f;
    if (k >= 0)
      d = sqdist(__ldg(xyz1 + j * 3 + 0) - __ldg(xyz2 + k * 3 + 0), __ldg(xyz1 + j * 3 + 1) - __ldg(xyz2 + k * 3 + 1),
                 __ldg(xyz1 + j * 3 + 2) - __ldg(xyz2 + k * 3 + 2));
    dist[j] = d;
  }
  cluster.sync();  // no CTA may exit while a peer can still read its shared memory (s_total)
}

// gradient for xyz1 only (emd_cuda.cu:284-300); one writer per element, so plain stores.
__global__ void __launch_bounds__(256)
emd_grad_kernel(long long total, int n, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                const float *__restrict__ graddist, const int *__restrict__ assignment,
                float *__restrict__ gradxyz1) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long cloud = i / n;
    const int j2 = __ldg(assignment + i);
    const float g = __ldg(graddist + i) * 2;
    if (j2 < 0 || j2 >= n) {  // unassigned (the forward leaves -1 when iters == 0; its dist is 0): zero gradient
      gradxyz1[i * 3 + 0] = gradxyz1[i * 3 + 1] = gradxyz1[i * 3 + 2] = 0.f;
      continue;
    }
    const long long o = (cloud * n + j2) * 3;
    gradxyz1[i * 3 + 0] = 0.f + g * (__ldg(xyz1 + i * 3 + 0) - __ldg(xyz2 + o + 0));
    gradxyz1[i * 3 + 1] = 0.f + g * (__ldg(xyz1 + i * 3 + 1) - __ldg(xyz2 + o + 1));
    gradxyz1[i * 3 + 2] = 0.f + g * (__ldg(xyz1 + i * 3 + 2) - __ldg(xyz2 + o + 2));
  }
}

static int emd_cluster_size(int b, int n) {
  int c = 8;
  while (c > 1 && ((long long)b * c > kNumSMs || n / c < 1024 || (n % c) != 0)) c >>= 1;
  return c;
}

static bool emd_smem_plan(int b, int n, int *cluster, int *tile_cap, size_t *smem) {
  const int c = emd_cluster_size(b, n);
  const size_t list_bytes = (size_t)(n / c) * 4;
  int cap = n < kEmdMaxTile ? n : kEmdMaxTile;
  const size_t limit = 220 * 1024;
  while (cap > 1024 && (size_t)cap * 16 + list_bytes > limit) cap -= 1024;
  if ((size_t)cap * 16 + list_bytes > limit) return false;
  *cluster = c;
  *tile_cap = cap;
  *smem = (size_t)cap * 16 + list_bytes;
  return true;
}

}  // namespace mvp

using namespace mvp;

MVP_API size_t mvp_emd_forward_workspace_bytes(int b, int n) {
  if (b <= 0 || n <= 0) return 16;
  return (size_t)b * n * 28;
}

// `granted`: the call site's per-device high-water mark of the kernel's dynamic shared memory (common.cuh)
template <typename K>
static int emd_launch(K kernel, size_t *granted, int b, int cluster, size_t smem, cudaStream_t stream, int n, int arg,
                      float eps, int iters, const float *xyz1, const float *xyz2, float *dist, int *assignment,
                      void *workspace) {
  {
    const int rc = grant_dyn_smem(kernel, smem, granted, 0);
    if (rc) return rc;
  }
  cudaError_t e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(b * cluster));
  cfg.blockDim = dim3(kEmdThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kernel, n, arg, eps, iters, xyz1, xyz2, dist, assignment, (unsigned char *)workspace);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return launch_status();
}

MVP_API int mvp_emd_forward_algo(int algo, int b, int n, int m, const float *xyz1, const float *xyz2, float eps,
                                 int iters, float *dist, int *assignment, void *workspace, size_t workspace_bytes,
                                 mvp_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || iters < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (algo < MVP_EMD_AUTO || algo > MVP_EMD_GRID) return MVP_ERR_INVALID_ARGUMENT;
  if (n != m) return MVP_ERR_EMD_SIZE_MISMATCH;      // emd_cuda.cu:236-239
  if (b > 512) return MVP_ERR_EMD_BATCH;             // emd_cuda.cu:241-244
  if (n % 1024 != 0) return MVP_ERR_EMD_MULTIPLE_1024;  // emd_cuda.cu:246-249
  if (b == 0 || n == 0) return MVP_OK;
  if (!xyz1 || !xyz2 || !dist || !assignment) return MVP_ERR_INVALID_ARGUMENT;
  if (!workspace || workspace_bytes < mvp_emd_forward_workspace_bytes(b, n)) return MVP_ERR_WORKSPACE;
  const bool grid_ok = n <= kEmdGridMaxN;
  if (algo == MVP_EMD_GRID && !grid_ok) return MVP_ERR_INVALID_ARGUMENT;
  if (grid_ok && algo != MVP_EMD_BRUTE) {
    const int cluster = emd_cluster_size(b, n);
    const int cap = std::max(8, n / kEmdGridPPC);
    const size_t smem = (size_t)n * 20 + sizeof(int) * (size_t)((cap + 1 + 3) & ~3) + sizeof(int) * (size_t)(n / cluster);
    static size_t granted_grid[kMaxDevices];
    return emd_launch(emd_auction_grid_kernel, granted_grid, b, cluster, smem, (cudaStream_t)stream, n, cap, eps, iters, xyz1, xyz2,
                      dist, assignment, workspace);
  }
  int cluster = 1, tile_cap = 0;
  size_t smem = 0;
  if (!emd_smem_plan(b, n, &cluster, &tile_cap, &smem)) return MVP_ERR_INVALID_ARGUMENT;
  static size_t granted_full[kMaxDevices];
  return emd_launch(emd_auction_kernel, granted_full, b, cluster, smem, (cudaStream_t)stream, n, tile_cap, eps, iters, xyz1, xyz2, dist,
                    assignment, workspace);
}

MVP_API int mvp_emd_forward(int b, int n, int m, const float *xyz1, const float *xyz2, float eps, int iters,
                            float *dist, int *assignment, void *workspace, size_t workspace_bytes,
                            mvp_stream_t stream) {
  return mvp_emd_forward_algo(MVP_EMD_AUTO, b, n, m, xyz1, xyz2, eps, iters, dist, assignment, workspace,
                              workspace_bytes, stream);
}

MVP_API int mvp_emd_backward(int b, int n, const float *xyz1, const float *xyz2, const float *graddist,
                             const int *assignment, float *gradxyz1, mvp_stream_t stream) {
  if (b < 0 || n < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0) return MVP_OK;
  if (!xyz1 || !xyz2 || !graddist || !assignment || !gradxyz1) return MVP_ERR_INVALID_ARGUMENT;
  const long long total = (long long)b * n;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
  emd_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(total, n, xyz1, xyz2, graddist, assignment,
                                                          gradxyz1);
  count_launch();
  return launch_status();
}
