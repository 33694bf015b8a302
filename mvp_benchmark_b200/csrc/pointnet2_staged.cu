// gather_points / group_points / three_interpolate and their backward scatters, shared-memory staged (sm_100a).
//
// Same results as the direct kernels in pointnet2.cu (which replace gather_points_cuda.cu:8-26,51-70,
// group_points_cuda.cu:10-31,56-79, three_interpolate_cuda.cu:11-35,61-84 of the reference) — these are the path taken
// when there are at least about as many gathered columns as source columns, which is every call the completion
// models make (M = 5 N, n = 2 m).
//
// Why: out[b,c,p] = points[b,c,idx[b,p]] reads 4 random bytes per output from a 12 KB row; every such load moves a
// 32-byte sector from L2, so the direct kernel is bound by L2 sector traffic at 8x the useful bytes (measured 24-36 %
// of the HBM copy peak in round 1).  Here a CTA owns G consecutive channel rows of one cloud — G*n contiguous floats —
// brings them into shared memory with ONE bulk TMA copy (cp.async.bulk, SASS UBLKCP), gathers from shared memory, and
// streams the outputs with coalesced stores; HBM/L2 see each source row once per column chunk.  The backward kernels
// accumulate into zeroed shared-memory rows with shared-memory atomics and write each gradient row once with plain
// stores: no global atomics, no memset.  (Accumulation order differs from the reference's global atomics, as theirs
// does from run to run; tests bound it at 1e-5 relative.)
#include "common.cuh"

namespace mvp {

constexpr int kStThreads = 512;
constexpr size_t kStRowBytes = 64 * 1024;  // shared-memory budget for the staged rows of a CTA (3 CTAs / SM)

// opt in to > 48 KB of dynamic shared memory, per kernel (TAG names the kernel: one high-water mark each) and device
template <int TAG, typename K>
static int set_smem(K kernel, size_t bytes) {
  static size_t granted[kMaxDevices];
  return grant_dyn_smem(kernel, bytes, granted);
}

__device__ __forceinline__ uint32_t st_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// rows[0 .. count) <- src[0 .. count): bulk TMA when 16-byte aligned, else plain loads.  Ends with a CTA barrier.
__device__ __forceinline__ void stage_rows(float *rows, const float *__restrict__ src, int count, uint64_t *bar) {
  const int tid = threadIdx.x;
  const bool bulk = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (count % 4 == 0);
  if (bulk) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_smem_u32(bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(count * 4) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       st_smem_u32(rows)),
                   "l"(src), "r"(count * 4), "r"(st_smem_u32(bar))
                   : "memory");
    }
    __syncthreads();  // the barrier is initialised before anybody polls it
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(st_smem_u32(bar))
        : "memory");
  } else {
    for (int i = tid; i < count; i += kStThreads) rows[i] = __ldg(src + i);
    __syncthreads();
  }
}

// grid (channel groups, column chunks, clouds)
__global__ void __launch_bounds__(kStThreads)
gather_staged_kernel(int c, int n, int mpts, int G, int chunk, const float *__restrict__ points,
                     const int *__restrict__ idx, float *__restrict__ out) {
  extern __shared__ __align__(128) float rows[];
  __shared__ __align__(8) uint64_t bar;
  const int b = blockIdx.z, c0 = blockIdx.x * G, gcount = min(G, c - c0);
  stage_rows(rows, points + ((size_t)b * c + c0) * n, gcount * n, &bar);
  const int p0 = blockIdx.y * chunk, p1 = min(mpts, p0 + chunk);
  const int *id = idx + (size_t)b * mpts;
  float *oo = out + ((size_t)b * c + c0) * mpts;
  for (int p = p0 + threadIdx.x; p < p1; p += kStThreads) {
    const int src = __ldg(id + p);
#pragma unroll 4
    for (int g = 0; g < gcount; g++) oo[(size_t)g * mpts + p] = rows[g * n + src];
  }
}

// grid (channel groups, 1, clouds): a CTA produces G complete gradient rows.  kVec: four consecutive columns per
// thread and 128-bit loads of idx / grad_out (rows 16-byte aligned, mpts % 4 == 0) — the loads of a warp are what keeps
// HBM busy while its lanes spin in the shared-memory CAS loops, so each one should carry as many bytes as possible.
template <bool kVec>
__global__ void __launch_bounds__(kStThreads)
gather_grad_staged_kernel(int c, int n, int mpts, int G, const float *__restrict__ grad_out,
                          const int *__restrict__ idx, float *__restrict__ grad_points) {
  extern __shared__ __align__(128) float rows[];
  const int b = blockIdx.z, c0 = blockIdx.x * G, gcount = min(G, c - c0);
  for (int i = threadIdx.x; i < gcount * n; i += kStThreads) rows[i] = 0.f;
  __syncthreads();
  const int *id = idx + (size_t)b * mpts;
  const float *go = grad_out + ((size_t)b * c + c0) * mpts;
  if (kVec) {
    for (int p = threadIdx.x * 4; p < mpts; p += kStThreads * 4) {
      const int4 dst = __ldg(reinterpret_cast<const int4 *>(id + p));
#pragma unroll 2
      for (int g = 0; g < gcount; g++) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(go + (size_t)g * mpts + p));
        float *r = rows + g * n;
        atomicAdd(r + dst.x, v.x);
        atomicAdd(r + dst.y, v.y);
        atomicAdd(r + dst.z, v.z);
        atomicAdd(r + dst.w, v.w);
      }
    }
  } else {
    for (int p = threadIdx.x; p < mpts; p += kStThreads) {
      const int dst = __ldg(id + p);
#pragma unroll 4
      for (int g = 0; g < gcount; g++) atomicAdd(&rows[g * n + dst], __ldg(go + (size_t)g * mpts + p));
    }
  }
  __syncthreads();
  float *gp = grad_points + ((size_t)b * c + c0) * n;
  for (int i = threadIdx.x; i < gcount * n; i += kStThreads) gp[i] = rows[i];
}

// out = fma(w2,p2, fma(w0,p0, w1*p1)) — the contraction nvcc gives three_interpolate_cuda.cu:33-34 (SASS-verified)
__global__ void __launch_bounds__(kStThreads)
three_interpolate_staged_kernel(int c, int m, int n, int G, int chunk, const float *__restrict__ points,
                                const int *__restrict__ idx, const float *__restrict__ weight,
                                float *__restrict__ out) {
  extern __shared__ __align__(128) float rows[];
  __shared__ __align__(8) uint64_t bar;
  const int b = blockIdx.z, c0 = blockIdx.x * G, gcount = min(G, c - c0);
  stage_rows(rows, points + ((size_t)b * c + c0) * m, gcount * m, &bar);
  const int p0 = blockIdx.y * chunk, p1 = min(n, p0 + chunk);
  float *oo = out + ((size_t)b * c + c0) * n;
  for (int p = p0 + threadIdx.x; p < p1; p += kStThreads) {
    const int *id = idx + ((size_t)b * n + p) * 3;
    const float *w = weight + ((size_t)b * n + p) * 3;
    const int i0 = __ldg(id + 0), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
    const float w0 = __ldg(w + 0), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
#pragma unroll 4
    for (int g = 0; g < gcount; g++) {
      const float *r = rows + g * m;
      oo[(size_t)g * n + p] = __fmaf_rn(w2, r[i2], __fmaf_rn(w0, r[i0], __fmul_rn(w1, r[i1])));
    }
  }
}

// kVec: four consecutive targets per thread — grad_out read as float4, their 12 indices and 12 weights as 3 + 3
// 128-bit loads (n % 4 == 0, 16-byte aligned rows).
template <bool kVec>
__global__ void __launch_bounds__(kStThreads)
three_interpolate_grad_staged_kernel(int c, int n, int m, int G, const float *__restrict__ grad_out,
                                     const int *__restrict__ idx, const float *__restrict__ weight,
                                     float *__restrict__ grad_points) {
  extern __shared__ __align__(128) float rows[];
  const int b = blockIdx.z, c0 = blockIdx.x * G, gcount = min(G, c - c0);
  for (int i = threadIdx.x; i < gcount * m; i += kStThreads) rows[i] = 0.f;
  __syncthreads();
  const float *go = grad_out + ((size_t)b * c + c0) * n;
  if (kVec) {
    for (int p = threadIdx.x * 4; p < n; p += kStThreads * 4) {
      int id[12];
      float w[12];
      const int4 *ip = reinterpret_cast<const int4 *>(idx + ((size_t)b * n + p) * 3);
      const float4 *wp = reinterpret_cast<const float4 *>(weight + ((size_t)b * n + p) * 3);
#pragma unroll
      for (int v = 0; v < 3; v++) {
        const int4 a = __ldg(ip + v);
        const float4 f = __ldg(wp + v);
        id[4 * v + 0] = a.x, id[4 * v + 1] = a.y, id[4 * v + 2] = a.z, id[4 * v + 3] = a.w;
        w[4 * v + 0] = f.x, w[4 * v + 1] = f.y, w[4 * v + 2] = f.z, w[4 * v + 3] = f.w;
      }
      for (int g = 0; g < gcount; g++) {
        const float4 gr4 = __ldg(reinterpret_cast<const float4 *>(go + (size_t)g * n + p));
        const float gr[4] = {gr4.x, gr4.y, gr4.z, gr4.w};
        float *r = rows + g * m;
#pragma unroll
        for (int e = 0; e < 12; e++) atomicAdd(r + id[e], __fmul_rn(gr[e / 3], w[e]));
      }
    }
  } else {
    for (int p = threadIdx.x; p < n; p += kStThreads) {
      const int *id = idx + ((size_t)b * n + p) * 3;
      const float *w = weight + ((size_t)b * n + p) * 3;
      const int i0 = __ldg(id + 0), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
      const float w0 = __ldg(w + 0), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
#pragma unroll 4
      for (int g = 0; g < gcount; g++) {
        const float gr = __ldg(go + (size_t)g * n + p);
        float *r = rows + g * m;
        atomicAdd(r + i0, __fmul_rn(gr, w0));
        atomicAdd(r + i1, __fmul_rn(gr, w1));
        atomicAdd(r + i2, __fmul_rn(gr, w2));
      }
    }
  }
  __syncthreads();
  float *gp = grad_points + ((size_t)b * c + c0) * m;
  for (int i = threadIdx.x; i < gcount * m; i += kStThreads) gp[i] = rows[i];
}

// ---- backward through a transposed index (CSR in a caller-provided workspace) -----------------------------------------
// grad_points[b,c,j] = sum over the gathered columns p with idx[b,p] == j of grad_out[b,c,p] (times a weight for
// three_interpolate).  The index is shared by all channels of a cloud, so it is transposed ONCE per cloud
// (scatter_csr_build_kernel: counting sort of the entries by destination with integer shared-memory atomics, which are
// native — fp32 ones are a CAS loop) into the workspace:  start[rows + 1] | perm[entries]  per cloud.  The consumers
// then bring G grad_out rows into shared memory with one bulk TMA copy and let every thread SUM the entries of its
// destinations and write the result with a coalesced store: no floating-point atomics, no memset, three CTAs per SM
// overlapping each other's copies.
constexpr int kCsrThreads = 1024;
constexpr int kCsrMaxRows = 48 * 1024;  // start[] lives in shared memory during the build (192 KB)

// per cloud: start[rows + 1] | perm[entries] (the gathered column of every entry, in destination order) |
// wperm[entries] (three_interpolate: the entry's weight, permuted alike — read coalesced by every channel group
// instead of gathered 4 bytes per 32-byte sector from weight[] each time)
static size_t csr_cloud_ints(int rows, int entries) { return ((size_t)rows + 1 + 2 * (size_t)entries + 3) & ~(size_t)3; }

// grid (clouds).  idx: `entries` destinations per cloud (entry e of gather: column e; of interpolate: (target e/3, k = e%3)).
__global__ void __launch_bounds__(kCsrThreads, 1)
scatter_csr_build_kernel(int rows, int entries, size_t cloud_ints, const int *__restrict__ idx,
                         const float *__restrict__ weight, int *__restrict__ ws) {
  extern __shared__ int start[];  // rows + 1
  __shared__ int s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int *id = idx + (size_t)blockIdx.x * entries;
  int *out_start = ws + (size_t)blockIdx.x * cloud_ints, *perm = out_start + rows + 1;
  for (int j = tid; j <= rows; j += kCsrThreads) start[j] = 0;
  __syncthreads();
  for (int e = tid; e < entries; e += kCsrThreads) atomicAdd(&start[__ldg(id + e) + 1], 1);
  __syncthreads();
  // exclusive scan of the counts: start[j + 1] := first slot of destination j (its cursor during the fill, after which
  // it has advanced to the first slot of destination j + 1); start[0] stays 0
  const int chunk = (rows + kCsrThreads - 1) / kCsrThreads;
  const int c0 = min(tid * chunk, rows), c1 = min(c0 + chunk, rows);
  int sum = 0;
  for (int c = c0; c < c1; c++) sum += start[c + 1];
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
  for (int c = c0; c < c1; c++) {
    const int cnt = start[c + 1];
    start[c + 1] = run;
    run += cnt;
  }
  __syncthreads();
  if (weight) {  // three_interpolate: entry e = 3 * target + k
    const float *w = weight + (size_t)blockIdx.x * entries;
    float *wperm = reinterpret_cast<float *>(perm + entries);
    for (int e = tid; e < entries; e += kCsrThreads) {
      const int slot = atomicAdd(&start[__ldg(id + e) + 1], 1);
      perm[slot] = e / 3;
      wperm[slot] = __ldg(w + e);
    }
  } else {
    for (int e = tid; e < entries; e += kCsrThreads) perm[atomicAdd(&start[__ldg(id + e) + 1], 1)] = e;
  }
  __syncthreads();
  for (int j = tid; j <= rows; j += kCsrThreads) out_start[j] = start[j];
}

// grid (channel groups, 1, clouds); dynamic shared memory: G rows of `cols` floats.
// kInterp: entry e = 3 * target + k carries the weight weight[b, e]; the row is indexed by the target.
template <bool kInterp, int G>
__global__ void __launch_bounds__(kStThreads)
scatter_csr_apply_kernel(int c, int rows, int cols, size_t cloud_ints, const float *__restrict__ grad_out,
                         const int *__restrict__ ws, float *__restrict__ grad_points) {
  extern __shared__ __align__(128) float rowbuf[];
  __shared__ __align__(8) uint64_t bar;
  const int b = blockIdx.z, c0 = blockIdx.x * G, gcount = min(G, c - c0);
  stage_rows(rowbuf, grad_out + ((size_t)b * c + c0) * cols, gcount * cols, &bar);
  const int *start = ws + (size_t)b * cloud_ints, *perm = start + rows + 1;
  const float *wperm = reinterpret_cast<const float *>(perm + (kInterp ? 3 * (size_t)cols : (size_t)cols));
  float *gp = grad_points + ((size_t)b * c + c0) * rows;
  for (int j = threadIdx.x; j < rows; j += kStThreads) {
    float acc[G];
#pragma unroll
    for (int g = 0; g < G; g++) acc[g] = 0.f;
    const int q1 = __ldg(start + j + 1);
    for (int q = __ldg(start + j); q < q1; q++) {
      const int p = __ldg(perm + q);
      const float wt = kInterp ? __ldg(wperm + q) : 1.f;
#pragma unroll
      for (int g = 0; g < G; g++)
        if (g < gcount) {
          const float v = rowbuf[g * cols + p];
          acc[g] = __fadd_rn(acc[g], kInterp ? __fmul_rn(v, wt) : v);  // product rounded first, as the atomics add it
        }
    }
#pragma unroll
    for (int g = 0; g < G; g++)
      if (g < gcount) gp[(size_t)g * rows + j] = acc[g];
  }
}

size_t scatter_csr_workspace_bytes(int b, int rows, int entries) {
  if (b <= 0 || rows <= 0 || entries <= 0 || rows > kCsrMaxRows) return 16;
  return sizeof(int) * (size_t)b * csr_cloud_ints(rows, entries);
}

// grad_points (b, c, rows) from grad_out (b, c, cols) through idx (b, entries): entries = cols (gather / group) or
// 3 * cols with weights (three_interpolate).  Returns -100 when this path does not apply (the caller falls back).
int scatter_csr_launch(bool interp, int b, int c, int rows, int cols, const float *grad_out, const int *idx,
                       const float *weight, float *grad_points, void *workspace, size_t workspace_bytes, cudaStream_t s) {
  const int entries = interp ? 3 * cols : cols;
  if (b > 65535 || rows > kCsrMaxRows) return -100;
  // Measured (tools/scatter_variants.py): the transposed index pays off when at least two grad_out rows fit the
  // shared-memory budget of a CTA (its index reads are then shared between channels) and most destinations receive
  // something; long rows and sparse scatters are faster in the workspace-free kernels.
  if ((size_t)cols * 4 * 2 > kStRowBytes || entries < rows) return -100;
  if (!workspace || workspace_bytes < scatter_csr_workspace_bytes(b, rows, entries)) return -100;
  const size_t cloud_ints = csr_cloud_ints(rows, entries);
  const size_t bsmem = sizeof(int) * ((size_t)rows + 1);
  {
    static size_t granted[kMaxDevices];
    const int rc0 = grant_dyn_smem(scatter_csr_build_kernel, bsmem > 40 * 1024 ? sizeof(int) * (kCsrMaxRows + 1) : bsmem, granted);
    if (rc0) return rc0;
  }
  scatter_csr_build_kernel<<<b, kCsrThreads, bsmem, s>>>(rows, entries, cloud_ints, idx, interp ? weight : nullptr,
                                                         (int *)workspace);
  // channels per CTA: as many as keep the staged rows within ~64 KB (three CTAs per SM)
  int G = 4;  // one of the instantiated values 4, 2, 1
  while (G > 1 && ((size_t)G * cols * 4 > kStRowBytes || G > c)) G >>= 1;
  const size_t smem = (size_t)G * cols * 4;
  dim3 grid((c + G - 1) / G, 1, b);
  int rc = MVP_OK;
#define MVP_CSR_APPLY(I, GG, TAG)                                                                                   \
  do {                                                                                                              \
    rc = set_smem<TAG>(scatter_csr_apply_kernel<I, GG>, smem);                                                      \
    if (!rc)                                                                                                        \
      scatter_csr_apply_kernel<I, GG><<<grid, kStThreads, smem, s>>>(c, rows, cols, cloud_ints, grad_out,            \
                                                                      (const int *)workspace, grad_points);         \
  } while (0)
  if (interp) {
    if (G >= 4) MVP_CSR_APPLY(true, 4, 10);
    else if (G >= 2) MVP_CSR_APPLY(true, 2, 11);
    else MVP_CSR_APPLY(true, 1, 12);
  } else {
    if (G >= 4) MVP_CSR_APPLY(false, 4, 13);
    else if (G >= 2) MVP_CSR_APPLY(false, 2, 14);
    else MVP_CSR_APPLY(false, 1, 15);
  }
#undef MVP_CSR_APPLY
  if (rc) return rc;
  count_launch(2);
  return launch_status();
}

// ---- gather + max over the neighbours (edge_preserve_sampling) ----------------------------------
// completion/model_utils.py:97-102 gathers the pk nearest neighbours' features of every sampled point with gather_points
// — a (B, C, M * pk) tensor, 250 MB at the first level — views it as (B, C, M, pk) and takes torch.max over pk.  Here
// the CTA that has a group of channel rows in shared memory takes the maximum over a point's pk neighbours on the spot:
// out[b,c,p] = max_j points[b,c,idx[b,p,j]], arg[b,c,p] = the neighbour index that attains it (the first j among equal
// maxima: torch.max's rule), which is all the backward pass needs — grad_points[b,c,arg[b,c,p]] += grad_out[b,c,p],
// accumulated in zeroed shared-memory rows and written once.  The (B, C, M, pk) tensor never exists.
constexpr int kGmG = 8;  // channel rows per CTA (register arrays of the running maxima)

__global__ void __launch_bounds__(kStThreads)
gather_max_staged_kernel(int c, int n, int mpts, int k, int G, int chunk, const float *__restrict__ points,
                         const int *__restrict__ idx, float *__restrict__ out, int *__restrict__ arg) {
  extern __shared__ __align__(128) float rows[];
  __shared__ __align__(8) uint64_t bar;
  const int b = blockIdx.z, c0 = blockIdx.x * G, gcount = min(G, c - c0);
  stage_rows(rows, points + ((size_t)b * c + c0) * n, gcount * n, &bar);
  const int p0 = blockIdx.y * chunk, p1 = min(mpts, p0 + chunk);
  float *oo = out + ((size_t)b * c + c0) * mpts;
  int *oa = arg + ((size_t)b * c + c0) * mpts;
  const float ninf = __int_as_float(0xff800000);
  for (int p = p0 + threadIdx.x; p < p1; p += kStThreads) {
    const int *id = idx + ((size_t)b * mpts + p) * k;
    float best[kGmG];
    int bi[kGmG];
#pragma unroll
    for (int g = 0; g < kGmG; g++) best[g] = ninf, bi[g] = 0;
    for (int j = 0; j < k; j++) {
      const int src = __ldg(id + j);
#pragma unroll
      for (int g = 0; g < kGmG; g++) {
        if (g < gcount) {
          const float v = rows[g * n + src];
          if (v > best[g] || j == 0) best[g] = v, bi[g] = src;  // strict: the first of equal maxima stays
        }
      }
    }
#pragma unroll
    for (int g = 0; g < kGmG; g++) {
      if (g < gcount) {
        oo[(size_t)g * mpts + p] = best[g];
        oa[(size_t)g * mpts + p] = bi[g];
      }
    }
  }
}

__global__ void __launch_bounds__(kStThreads)
gather_max_grad_staged_kernel(int c, int n, int mpts, int G, const float *__restrict__ grad_out,
                              const int *__restrict__ arg, float *__restrict__ grad_points) {
  extern __shared__ __align__(128) float rows[];
  const int b = blockIdx.z, c0 = blockIdx.x * G, gcount = min(G, c - c0);
  for (int i = threadIdx.x; i < gcount * n; i += kStThreads) rows[i] = 0.f;
  __syncthreads();
  const float *go = grad_out + ((size_t)b * c + c0) * mpts;
  const int *ga = arg + ((size_t)b * c + c0) * mpts;
  for (int g = 0; g < gcount; g++)
    for (int p = threadIdx.x; p < mpts; p += kStThreads)
      atomicAdd(&rows[g * n + __ldg(ga + (size_t)g * mpts + p)], __ldg(go + (size_t)g * mpts + p));
  __syncthreads();
  float *gp = grad_points + ((size_t)b * c + c0) * n;
  for (int i = threadIdx.x; i < gcount * n; i += kStThreads) gp[i] = rows[i];
}

// ---- attention-weighted sum over the neighbours (SA_module, vrcnet.py:49-52) ---------------
// After the 1x1 convolutions have been moved in front of the gather (model_patches.sa_module_forward), SA_module still
// materialises x3 = get_edge_features(y3, idx) (B, C, k, N), repeats the attention weights w (B, C/S, k, N) S times
// along the channels, multiplies and sums over k: four (B, C, k, N) tensors forward, as many backward.  Here
//     out[b,c,p] = sum_j w[b, c mod Cw, j, p] * y[b, c, idx[b,p,j]]
// is one kernel: a CTA owns the S channels { cw, cw + Cw, ... } that SHARE a weight row, stages their S rows of y in
// shared memory, and a thread walks the k neighbours of its point once — one index load, one weight load (coalesced
// along p), S shared-memory reads and S fmas per neighbour.  Backward: grad_w[b,cw,j,p] = sum_s g[b,c_s,p] y[b,c_s,idx]
// (same staging, summed in registers) and grad_y[b,c,m] = sum over { (p,j) : idx[b,p,j] = m } of g[b,c,p] w[b,cw,j,p]
// (accumulated in zeroed shared-memory rows, written once: no global atomics).
constexpr int kNwS = 8;  // channels sharing a weight row (share_planes)

// MODE 0: forward (out).  MODE 1: grad_w (a = grad_out).  grid (Cw, column chunks, b)
template <int MODE>
__global__ void __launch_bounds__(kStThreads)
nbr_wsum_kernel(int c, int cw, int n, int k, int S, int chunk, const float *__restrict__ y, const int *__restrict__ idx,
                const float *__restrict__ w, const float *__restrict__ a, float *__restrict__ out) {
  extern __shared__ __align__(128) float rows[];
  const int b = blockIdx.z, w0 = blockIdx.x;
  for (int s_ = 0; s_ < S; s_++) {
    const float *src = y + ((size_t)b * c + w0 + (size_t)s_ * cw) * n;
    for (int i = threadIdx.x; i < n; i += kStThreads) rows[s_ * n + i] = __ldg(src + i);
  }
  __syncthreads();
  const int p0 = blockIdx.y * chunk, p1 = min(n, p0 + chunk);
  for (int p = p0 + threadIdx.x; p < p1; p += kStThreads) {
    const int *id = idx + ((size_t)b * n + p) * k;
    const float *wp = w + ((size_t)b * cw + w0) * k * n + p;
    float acc[kNwS];
#pragma unroll
    for (int s_ = 0; s_ < kNwS; s_++) {
      acc[s_] = 0.f;
      if (MODE == 1 && s_ < S) acc[s_] = __ldg(a + ((size_t)b * c + w0 + (size_t)s_ * cw) * n + p);  // upstream gradients
    }
    for (int j = 0; j < k; j++) {
      const int src = __ldg(id + j);
      if (MODE == 0) {
        const float wv = __ldg(wp + (size_t)j * n);
#pragma unroll
        for (int s_ = 0; s_ < kNwS; s_++)
          if (s_ < S) acc[s_] = __fmaf_rn(wv, rows[s_ * n + src], acc[s_]);
      } else {
        float d = 0.f;
#pragma unroll
        for (int s_ = 0; s_ < kNwS; s_++)
          if (s_ < S) d = __fmaf_rn(acc[s_], rows[s_ * n + src], d);
        out[(((size_t)b * cw + w0) * k + j) * n + p] = d;
      }
    }
    if (MODE == 0) {
#pragma unroll
      for (int s_ = 0; s_ < kNwS; s_++)
        if (s_ < S) out[((size_t)b * c + w0 + (size_t)s_ * cw) * n + p] = acc[s_];
    }
  }
}

// grad_y: grid (Cw, channel sub-groups, b): a CTA produces the complete gradient rows of `per` of the S channels that
// share weight row w0 (all S when there are enough CTAs without splitting)
__global__ void __launch_bounds__(kStThreads)
nbr_wsum_grad_y_kernel(int c, int cw, int n, int k, int S_all, int per, const int *__restrict__ idx,
                       const float *__restrict__ w, const float *__restrict__ g, float *__restrict__ grad_y) {
  extern __shared__ __align__(128) float rows[];
  const int b = blockIdx.z, s_lo = blockIdx.y * per, S = min(per, S_all - s_lo);
  const int w0 = blockIdx.x + s_lo * cw;  // first channel of this CTA; the others follow at a stride of cw
  if (S <= 0) return;
  for (int i = threadIdx.x; i < S * n; i += kStThreads) rows[i] = 0.f;
  __syncthreads();
  for (int p = threadIdx.x; p < n; p += kStThreads) {
    const int *id = idx + ((size_t)b * n + p) * k;
    const float *wp = w + ((size_t)b * cw + blockIdx.x) * k * n + p;
    float gv[kNwS];
#pragma unroll
    for (int s_ = 0; s_ < kNwS; s_++) gv[s_] = s_ < S ? __ldg(g + ((size_t)b * c + w0 + (size_t)s_ * cw) * n + p) : 0.f;
    for (int j = 0; j < k; j++) {
      const int src = __ldg(id + j);
      const float wv = __ldg(wp + (size_t)j * n);
#pragma unroll
      for (int s_ = 0; s_ < kNwS; s_++)
        if (s_ < S) atomicAdd(&rows[s_ * n + src], gv[s_] * wv);
    }
  }
  __syncthreads();
  for (int s_ = 0; s_ < S; s_++) {
    float *dst = grad_y + ((size_t)b * c + w0 + (size_t)s_ * cw) * n;
    for (int i = threadIdx.x; i < n; i += kStThreads) dst[i] = rows[s_ * n + i];
  }
}

// ---- launch plans --------------------------------------------------------------------------------------------------
// rows: source columns per channel (n for gather, m for interpolate); cols: gathered columns per channel
bool staged_applicable(int b, int c, int rows, int cols) {
  return b > 0 && b <= 65535 && c > 0 && rows > 0 && (size_t)rows * 4 <= 200 * 1024 && (long long)cols * 2 >= rows;
}

static int group_size(int c, int rows) {
  int G = 8;
  while (G > 1 && (size_t)G * rows * 4 > kStRowBytes) G >>= 1;
  return std::min(G, c);
}

// column chunks: enough CTAs for ~2 waves, but never chunks so small that re-staging the rows dominates
static int column_chunk(int b, int groups, int rows, int cols) {
  int split = 1;
  while ((long long)b * groups * split < 2LL * 3 * kNumSMs && cols / (split * 2) >= 4 * rows) split *= 2;
  const int chunk = (cols + split - 1) / split;
  return (chunk + kStThreads - 1) / kStThreads * kStThreads;
}

int gather_staged_launch(int b, int c, int n, int mpts, const float *points, const int *idx, float *out,
                         cudaStream_t s) {
  const int G = group_size(c, n), groups = (c + G - 1) / G;
  const int chunk = column_chunk(b, groups, n, mpts);
  const size_t smem = (size_t)G * n * 4;
  if (int rc = set_smem<0>(gather_staged_kernel, smem)) return rc;
  dim3 grid(groups, (mpts + chunk - 1) / chunk, b);
  gather_staged_kernel<<<grid, kStThreads, smem, s>>>(c, n, mpts, G, chunk, points, idx, out);
  count_launch();
  return launch_status();
}

int gather_grad_staged_launch(int b, int c, int n, int mpts, const float *grad_out, const int *idx, float *grad_points,
                              cudaStream_t s) {
  const int G = group_size(c, n), groups = (c + G - 1) / G;
  const size_t smem = (size_t)G * n * 4;
  const bool vec = (mpts % 4 == 0) && ((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(idx)) & 15) == 0;
  if (vec) {
    if (int rc = set_smem<1>(gather_grad_staged_kernel<true>, smem)) return rc;
    gather_grad_staged_kernel<true><<<dim3(groups, 1, b), kStThreads, smem, s>>>(c, n, mpts, G, grad_out, idx, grad_points);
  } else {
    if (int rc = set_smem<6>(gather_grad_staged_kernel<false>, smem)) return rc;
    gather_grad_staged_kernel<false><<<dim3(groups, 1, b), kStThreads, smem, s>>>(c, n, mpts, G, grad_out, idx, grad_points);
  }
  count_launch();
  return launch_status();
}

int three_interpolate_staged_launch(int b, int c, int m, int n, const float *points, const int *idx, const float *weight,
                                    float *out, cudaStream_t s) {
  const int G = group_size(c, m), groups = (c + G - 1) / G;
  const int chunk = column_chunk(b, groups, m, n);
  const size_t smem = (size_t)G * m * 4;
  if (int rc = set_smem<2>(three_interpolate_staged_kernel, smem)) return rc;
  dim3 grid(groups, (n + chunk - 1) / chunk, b);
  three_interpolate_staged_kernel<<<grid, kStThreads, smem, s>>>(c, m, n, G, chunk, points, idx, weight, out);
  count_launch();
  return launch_status();
}

int three_interpolate_grad_staged_launch(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                         const float *weight, float *grad_points, cudaStream_t s) {
  const int G = group_size(c, m), groups = (c + G - 1) / G;
  const size_t smem = (size_t)G * m * 4;
  const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(idx) |
                                     reinterpret_cast<uintptr_t>(weight)) & 15) == 0;
  if (vec) {
    if (int rc = set_smem<3>(three_interpolate_grad_staged_kernel<true>, smem)) return rc;
    three_interpolate_grad_staged_kernel<true><<<dim3(groups, 1, b), kStThreads, smem, s>>>(c, n, m, G, grad_out, idx,
                                                                                          weight, grad_points);
  } else {
    if (int rc = set_smem<7>(three_interpolate_grad_staged_kernel<false>, smem)) return rc;
    three_interpolate_grad_staged_kernel<false><<<dim3(groups, 1, b), kStThreads, smem, s>>>(c, n, m, G, grad_out, idx,
                                                                                           weight, grad_points);
  }
  count_launch();
  return launch_status();
}


int gather_max_launch(int b, int c, int n, int mpts, int k, const float *points, const int *idx, float *out, int *arg,
                      cudaStream_t s) {
  const int G = std::min(group_size(c, n), kGmG), groups = (c + G - 1) / G;
  const int chunk = column_chunk(b, groups, n, mpts * std::max(k, 1));  // (k reads per output: more columns' worth of work)
  const size_t smem = (size_t)G * n * 4;
  if (int rc = set_smem<20>(gather_max_staged_kernel, smem)) return rc;
  dim3 grid(groups, (mpts + chunk - 1) / chunk, b);
  gather_max_staged_kernel<<<grid, kStThreads, smem, s>>>(c, n, mpts, k, G, chunk, points, idx, out, arg);
  count_launch();
  return launch_status();
}

int gather_max_grad_launch(int b, int c, int n, int mpts, const float *grad_out, const int *arg, float *grad_points,
                           cudaStream_t s) {
  const int G = std::min(group_size(c, n), kGmG), groups = (c + G - 1) / G;
  const size_t smem = (size_t)G * n * 4;
  if (int rc = set_smem<21>(gather_max_grad_staged_kernel, smem)) return rc;
  gather_max_grad_staged_kernel<<<dim3(groups, 1, b), kStThreads, smem, s>>>(c, n, mpts, G, grad_out, arg, grad_points);
  count_launch();
  return launch_status();
}


int nbr_wsum_launch(int mode, int b, int c, int cw, int n, int k, const float *y, const int *idx, const float *w,
                    const float *a, float *out, cudaStream_t s) {
  const int S = c / cw;
  const size_t smem = (size_t)S * n * 4;
  if (mode == 2) {
    int per = S;  // channels per CTA: halved until the grid fills the machine about twice
    while (per > 1 && (long long)b * cw * ((S + per - 1) / per) < 2LL * kNumSMs) per = (per + 1) / 2;
    if (int rc = set_smem<32>(nbr_wsum_grad_y_kernel, smem)) return rc;
    nbr_wsum_grad_y_kernel<<<dim3(cw, (S + per - 1) / per, b), kStThreads, (size_t)per * n * 4, s>>>(c, cw, n, k, S, per, idx, w,
                                                                                                a, out);
  } else {
    int split = 1;
    while ((long long)b * cw * split < 2LL * kNumSMs && n / (split * 2) >= kStThreads) split *= 2;
    const int chunk = ((n + split - 1) / split + kStThreads - 1) / kStThreads * kStThreads;
    dim3 grid(cw, (n + chunk - 1) / chunk, b);
    if (mode == 0) {
      if (int rc = set_smem<30>(nbr_wsum_kernel<0>, smem)) return rc;
      nbr_wsum_kernel<0><<<grid, kStThreads, smem, s>>>(c, cw, n, k, S, chunk, y, idx, w, nullptr, out);
    } else {
      if (int rc = set_smem<31>(nbr_wsum_kernel<1>, smem)) return rc;
      nbr_wsum_kernel<1><<<grid, kStThreads, smem, s>>>(c, cw, n, k, S, chunk, y, idx, w, a, out);
    }
  }
  count_launch();
  return launch_status();
}

}  // namespace mvp

// gather_points over the pk neighbours of every sampled point followed by a maximum over the neighbours
// (completion/model_utils.py:97-102) as ONE launch.  points (b,c,n), idx (b,npoints,k) -> out (b,c,npoints) and
// arg (b,c,npoints), the neighbour index that attains each maximum (the first among equals).  n * 4 bytes <= 64 KB.
MVP_API int mvp_gather_max(int b, int c, int n, int npoints, int k, const float *points, const int *idx, float *out,
                           int *arg, mvp_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || npoints < 0 || k <= 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || c == 0 || npoints == 0) return MVP_OK;
  if (!points || !idx || !out || !arg || b > 65535 || (size_t)n * 4 > mvp::kStRowBytes) return MVP_ERR_INVALID_ARGUMENT;
  return mvp::gather_max_launch(b, c, n, npoints, k, points, idx, out, arg, (cudaStream_t)stream);
}

// its backward: grad_points (b,c,n), fully written, = grad_out (b,c,npoints) scattered by arg (b,c,npoints)
MVP_API int mvp_gather_max_grad(int b, int c, int n, int npoints, const float *grad_out, const int *arg,
                                float *grad_points, mvp_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || npoints < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || c == 0) return MVP_OK;
  if (!grad_out || !arg || !grad_points || b > 65535 || (size_t)n * 4 > mvp::kStRowBytes) return MVP_ERR_INVALID_ARGUMENT;
  return mvp::gather_max_grad_launch(b, c, n, npoints, grad_out, arg, grad_points, (cudaStream_t)stream);
}

// out[b,ch,p] = sum_j w[b, ch mod cw, j, p] * y[b, ch, idx[b,p,j]]  (SA_module's weighted neighbour aggregation,
// completion/models/vrcnet.py:49-52, without its (B, C, k, N) intermediates).  y (b,c,n), idx (b,n,k) int32,
// w (b,cw,k,n), c = S * cw with S <= 8, n <= 6144.
static bool nbr_wsum_ok(int b, int c, int cw, int n, int k) {
  return b > 0 && b <= 65535 && c > 0 && cw > 0 && c % cw == 0 && c / cw <= mvp::kNwS && n > 0 && k > 0 &&
         (size_t)(c / cw) * n * 4 <= 200 * 1024;
}
MVP_API int mvp_neighbor_weighted_sum(int b, int c, int cw, int n, int k, const float *y, const int *idx, const float *w,
                                      float *out, mvp_stream_t stream) {
  if (b == 0) return MVP_OK;
  if (!nbr_wsum_ok(b, c, cw, n, k) || !y || !idx || !w || !out) return MVP_ERR_INVALID_ARGUMENT;
  return mvp::nbr_wsum_launch(0, b, c, cw, n, k, y, idx, w, nullptr, out, (cudaStream_t)stream);
}
// its backward: grad_w (b,cw,k,n) and grad_y (b,c,n), both fully written, from grad_out (b,c,n)
MVP_API int mvp_neighbor_weighted_sum_grad(int b, int c, int cw, int n, int k, const float *y, const int *idx,
                                           const float *w, const float *grad_out, float *grad_y, float *grad_w,
                                           mvp_stream_t stream) {
  if (b == 0) return MVP_OK;
  if (!nbr_wsum_ok(b, c, cw, n, k) || !y || !idx || !w || !grad_out || !grad_y || !grad_w) return MVP_ERR_INVALID_ARGUMENT;
  int rc = mvp::nbr_wsum_launch(1, b, c, cw, n, k, y, idx, w, grad_out, grad_w, (cudaStream_t)stream);
  if (rc) return rc;
  return mvp::nbr_wsum_launch(2, b, c, cw, n, k, y, idx, w, grad_out, grad_y, (cudaStream_t)stream);
}
