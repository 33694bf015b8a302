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

// grid (channel groups, 1, clouds): a CTA produces G complete gradient rows
__global__ void __launch_bounds__(kStThreads)
gather_grad_staged_kernel(int c, int n, int mpts, int G, const float *__restrict__ grad_out,
                          const int *__restrict__ idx, float *__restrict__ grad_points) {
  extern __shared__ __align__(128) float rows[];
  const int b = blockIdx.z, c0 = blockIdx.x * G, gcount = min(G, c - c0);
  for (int i = threadIdx.x; i < gcount * n; i += kStThreads) rows[i] = 0.f;
  __syncthreads();
  const int *id = idx + (size_t)b * mpts;
  const float *go = grad_out + ((size_t)b * c + c0) * mpts;
  for (int p = threadIdx.x; p < mpts; p += kStThreads) {
    const int dst = __ldg(id + p);
#pragma unroll 4
    for (int g = 0; g < gcount; g++) atomicAdd(&rows[g * n + dst], __ldg(go + (size_t)g * mpts + p));
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

__global__ void __launch_bounds__(kStThreads)
three_interpolate_grad_staged_kernel(int c, int n, int m, int G, const float *__restrict__ grad_out,
                                     const int *__restrict__ idx, const float *__restrict__ weight,
                                     float *__restrict__ grad_points) {
  extern __shared__ __align__(128) float rows[];
  const int b = blockIdx.z, c0 = blockIdx.x * G, gcount = min(G, c - c0);
  for (int i = threadIdx.x; i < gcount * m; i += kStThreads) rows[i] = 0.f;
  __syncthreads();
  const float *go = grad_out + ((size_t)b * c + c0) * n;
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
  __syncthreads();
  float *gp = grad_points + ((size_t)b * c + c0) * m;
  for (int i = threadIdx.x; i < gcount * m; i += kStThreads) gp[i] = rows[i];
}

// ---- backward through a transposed index (CSR) ---------------------------------------------------------------------
// grad_points[b,c,j] = sum over the gathered columns p with idx[b,p] == j of grad_out[b,c,p].  The index is shared by
// all channels of a cloud, so a CTA transposes it ONCE in shared memory (counting sort of the columns by destination:
// integer shared-memory atomics, which are native — fp32 ones are a CAS loop) and then, channel after channel, brings
// the grad_out row in by bulk TMA (double-buffered) and lets each thread SUM the columns of its destinations and write
// the result with a coalesced store.  No floating-point atomics at all, no memset.
__device__ __forceinline__ void csr_tma_issue(float *dst, const float *src, int count, uint64_t *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(count * 4) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   st_smem_u32(dst)),
               "l"(src), "r"(count * 4), "r"(st_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void csr_bar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(st_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Exclusive scan of start[1 .. rows] in place (start[0] = 0 is kept): start[j + 1] := number of entries with
// destination < j + 1 ... i.e. after the call start[j] is the first slot of destination j and start[rows] the total.
// Done in two steps so that start[j + 1] can serve as destination j's fill cursor: see the callers.
__device__ __forceinline__ void csr_scan(int *start, int rows, int *s_warp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = (rows + kStThreads - 1) / kStThreads;
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
    int v = lane < kStThreads / 32 ? s_warp[lane] : 0;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, off);
      if (lane >= off) v += u;
    }
    s_warp[lane] = v;
  }
  __syncthreads();
  int run = incl - sum + (warp ? s_warp[warp - 1] : 0);
  for (int c = c0; c < c1; c++) {  // start[c + 1] := first slot of destination c (its cursor during the fill)
    const int cnt = start[c + 1];
    start[c + 1] = run;
    run += cnt;
  }
  __syncthreads();
}

// grid (channel groups, 1, clouds).  Shared memory: start[n + 1 (+pad)] | perm[mpts] | two row buffers of mpts floats.
__global__ void __launch_bounds__(kStThreads)
gather_grad_csr_kernel(int c, int n, int mpts, int G, const float *__restrict__ grad_out, const int *__restrict__ idx,
                       float *__restrict__ grad_points) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int s_warp[32];
  const int mp4 = (mpts + 3) & ~3;
  float *const row0 = reinterpret_cast<float *>(smem_raw), *const row1 = row0 + mp4;
  int *perm = reinterpret_cast<int *>(smem_raw) + 2 * mp4;
  int *start = perm + mp4;  // n + 1
  const int tid = threadIdx.x;
  const int b = blockIdx.z, c0 = blockIdx.x * G, gcount = min(G, c - c0);
  const int *id = idx + (size_t)b * mpts;
  const float *go = grad_out + ((size_t)b * c + c0) * mpts;
  const bool bulk = ((reinterpret_cast<uintptr_t>(go) & 15) == 0) && (mpts % 4 == 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (bulk) csr_tma_issue(row0, go, mpts, &bar[0]);  // the first row streams in while the index is transposed
  }
  for (int j = tid; j <= n; j += kStThreads) start[j] = 0;
  __syncthreads();
  for (int p = tid; p < mpts; p += kStThreads) atomicAdd(&start[__ldg(id + p) + 1], 1);
  __syncthreads();
  csr_scan(start, n, s_warp);
  for (int p = tid; p < mpts; p += kStThreads) perm[atomicAdd(&start[__ldg(id + p) + 1], 1)] = p;
  __syncthreads();  // start[j] = first slot of destination j, start[n] = mpts

  float *gp = grad_points + ((size_t)b * c + c0) * n;
  for (int g = 0; g < gcount; g++) {
    float *r = (g & 1) ? row1 : row0;
    if (bulk) {
      csr_bar_wait(&bar[g & 1], (g >> 1) & 1);
      if (tid == 0 && g + 1 < gcount) csr_tma_issue((g & 1) ? row0 : row1, go + (size_t)(g + 1) * mpts, mpts, &bar[(g + 1) & 1]);
    } else {
      for (int p = tid; p < mpts; p += kStThreads) r[p] = __ldg(go + (size_t)g * mpts + p);
      __syncthreads();
    }
    for (int j = tid; j < n; j += kStThreads) {
      float acc = 0.f;
      for (int q = start[j]; q < start[j + 1]; q++) acc = __fadd_rn(acc, r[perm[q]]);
      gp[(size_t)g * n + j] = acc;
    }
    __syncthreads();  // everybody is done with row[g & 1] before it is refilled two iterations later
  }
}

// three_interpolate backward: grad_points[b,c,j] = sum over (p,k) with idx[b,p,k] == j of grad_out[b,c,p] * w[b,p,k]
// (products rounded before the sum, as three_interpolate_cuda.cu:81-83 adds them).
// Shared memory: two row buffers of n floats | permp[3n] | permw[3n] | start[m + 1].
__global__ void __launch_bounds__(kStThreads)
three_interpolate_grad_csr_kernel(int c, int n, int m, int G, const float *__restrict__ grad_out,
                                  const int *__restrict__ idx, const float *__restrict__ weight,
                                  float *__restrict__ grad_points) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int s_warp[32];
  const int n4 = (n + 3) & ~3;
  float *const row0 = reinterpret_cast<float *>(smem_raw), *const row1 = row0 + n4;
  int *permp = reinterpret_cast<int *>(smem_raw) + 2 * n4;
  float *permw = reinterpret_cast<float *>(permp + 3 * n);
  int *start = reinterpret_cast<int *>(permw + 3 * n);  // m + 1
  const int tid = threadIdx.x;
  const int b = blockIdx.z, c0 = blockIdx.x * G, gcount = min(G, c - c0);
  const int *id = idx + (size_t)b * n * 3;
  const float *w = weight + (size_t)b * n * 3;
  const float *go = grad_out + ((size_t)b * c + c0) * n;
  const bool bulk = ((reinterpret_cast<uintptr_t>(go) & 15) == 0) && (n % 4 == 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (bulk) csr_tma_issue(row0, go, n, &bar[0]);
  }
  for (int j = tid; j <= m; j += kStThreads) start[j] = 0;
  __syncthreads();
  for (int e = tid; e < 3 * n; e += kStThreads) atomicAdd(&start[__ldg(id + e) + 1], 1);
  __syncthreads();
  csr_scan(start, m, s_warp);
  for (int e = tid; e < 3 * n; e += kStThreads) {
    const int slot = atomicAdd(&start[__ldg(id + e) + 1], 1);
    permp[slot] = e / 3;
    permw[slot] = __ldg(w + e);
  }
  __syncthreads();

  float *gp = grad_points + ((size_t)b * c + c0) * m;
  for (int g = 0; g < gcount; g++) {
    float *r = (g & 1) ? row1 : row0;
    if (bulk) {
      csr_bar_wait(&bar[g & 1], (g >> 1) & 1);
      if (tid == 0 && g + 1 < gcount) csr_tma_issue((g & 1) ? row0 : row1, go + (size_t)(g + 1) * n, n, &bar[(g + 1) & 1]);
    } else {
      for (int p = tid; p < n; p += kStThreads) r[p] = __ldg(go + (size_t)g * n + p);
      __syncthreads();
    }
    for (int j = tid; j < m; j += kStThreads) {
      float acc = 0.f;
      for (int q = start[j]; q < start[j + 1]; q++) acc = __fadd_rn(acc, __fmul_rn(r[permp[q]], permw[q]));
      gp[(size_t)g * m + j] = acc;
    }
    __syncthreads();
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

// opt in to > 48 KB of dynamic shared memory, once per kernel (TAG names the kernel: one high-water mark each) and size
template <int TAG, typename K>
static int set_smem(K kernel, size_t bytes) {
  static size_t granted = 40 * 1024;  // static + dynamic share the default 48 KB: opt in a little below it
  if (bytes <= granted) return MVP_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return (int)e;
  granted = bytes;
  return MVP_OK;
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

// channels per CTA of the CSR kernels: amortise the transposition over several rows, but keep >= ~2 waves of CTAs
static int csr_group(int b, int c, int ctas_per_sm) {
  int G = 16;
  while (G > 2 && (long long)b * ((c + G - 1) / G) < 2LL * ctas_per_sm * kNumSMs) G >>= 1;
  return std::min(G, c);
}

int gather_grad_staged_launch(int b, int c, int n, int mpts, const float *grad_out, const int *idx, float *grad_points,
                              cudaStream_t s) {
  const size_t csr_smem = sizeof(float) * (3 * (size_t)((mpts + 3) & ~3) + n + 1);
  if (csr_smem <= 200 * 1024) {
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / csr_smem));
    const int G = csr_group(b, c, per_sm);
    if (int rc = set_smem<4>(gather_grad_csr_kernel, csr_smem)) return rc;
    gather_grad_csr_kernel<<<dim3((c + G - 1) / G, 1, b), kStThreads, csr_smem, s>>>(c, n, mpts, G, grad_out, idx,
                                                                                    grad_points);
    count_launch();
    return launch_status();
  }
  const int G = group_size(c, n), groups = (c + G - 1) / G;
  const size_t smem = (size_t)G * n * 4;
  if (int rc = set_smem<1>(gather_grad_staged_kernel, smem)) return rc;
  gather_grad_staged_kernel<<<dim3(groups, 1, b), kStThreads, smem, s>>>(c, n, mpts, G, grad_out, idx, grad_points);
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
  const size_t csr_smem = sizeof(float) * (2 * (size_t)((n + 3) & ~3) + 6 * (size_t)n + m + 1);
  if (csr_smem <= 200 * 1024) {
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / csr_smem));
    const int G = csr_group(b, c, per_sm);
    if (int rc = set_smem<5>(three_interpolate_grad_csr_kernel, csr_smem)) return rc;
    three_interpolate_grad_csr_kernel<<<dim3((c + G - 1) / G, 1, b), kStThreads, csr_smem, s>>>(c, n, m, G, grad_out, idx,
                                                                                               weight, grad_points);
    count_launch();
    return launch_status();
  }
  const int G = group_size(c, m), groups = (c + G - 1) / G;
  const size_t smem = (size_t)G * m * 4;
  if (int rc = set_smem<3>(three_interpolate_grad_staged_kernel, smem)) return rc;
  three_interpolate_grad_staged_kernel<<<dim3(groups, 1, b), kStThreads, smem, s>>>(c, n, m, G, grad_out, idx, weight,
                                                                                  grad_points);
  count_launch();
  return launch_status();
}

}  // namespace mvp
