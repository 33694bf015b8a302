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

// gradient for xyz1 only (emd_cuda.cu:284-300); one writer per element, so plain stores.
__global__ void __launch_bounds__(256)
emd_grad_kernel(long long total, int n, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                const float *__restrict__ graddist, const int *__restrict__ assignment,
                float *__restrict__ gradxyz1) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long cloud = i / n;
    const int j2 = __ldg(assignment + i);
    const float g = __ldg(graddist + i) * 2;
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

MVP_API int mvp_emd_forward(int b, int n, int m, const float *xyz1, const float *xyz2, float eps, int iters,
                            float *dist, int *assignment, void *workspace, size_t workspace_bytes,
                            mvp_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || iters < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (n != m) return MVP_ERR_EMD_SIZE_MISMATCH;      // emd_cuda.cu:236-239
  if (b > 512) return MVP_ERR_EMD_BATCH;             // emd_cuda.cu:241-244
  if (n % 1024 != 0) return MVP_ERR_EMD_MULTIPLE_1024;  // emd_cuda.cu:246-249
  if (b == 0 || n == 0) return MVP_OK;
  if (!xyz1 || !xyz2 || !dist || !assignment) return MVP_ERR_INVALID_ARGUMENT;
  if (!workspace || workspace_bytes < mvp_emd_forward_workspace_bytes(b, n)) return MVP_ERR_WORKSPACE;
  int cluster = 1, tile_cap = 0;
  size_t smem = 0;
  if (!emd_smem_plan(b, n, &cluster, &tile_cap, &smem)) return MVP_ERR_INVALID_ARGUMENT;
  cudaError_t e = cudaFuncSetAttribute(emd_auction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(b * cluster));
  cfg.blockDim = dim3(kEmdThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, emd_auction_kernel, n, tile_cap, eps, iters, xyz1, xyz2, dist, assignment,
                         (unsigned char *)workspace);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return launch_status();
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
