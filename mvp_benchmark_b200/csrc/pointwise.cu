// Shared 1x1 convolution over a point cloud's feature map — the contraction SURVEY.md §8f row 4 names ("the
// grouped-feature MLP folded in as a dense contraction"): y[b, co, p] = sum_ci w[co, ci] * x[b, ci, p] + bias[co],
// optionally through ReLU, on the 5th-generation tensor cores.  The callers are the nn.Conv1d / nn.Conv2d(kernel 1)
// layers of completion/models/{pcn,ecg,vrcnet}.py and completion/model_utils.py (after model_patches moves them in
// front of the neighbour gather, every one of them is this plain per-point contraction).
//
// What is in this file:
//   pointwise_conv_kernel       both operands streamed through two stages — any shape, and the masked form (input
//                               gradient behind a ReLU)
//   pointwise_resident_kernel   the HBM-bound layers: weight tile resident in shared memory, persistent CTAs, two
//                               accumulators in tensor memory (the one that is fast: 66 % of the HBM peak at 64 -> 256)
//   pointwise_wgrad_kernel      weight + bias gradient: K = points, both operands K-major as they lie in HBM
//   bias_add_kernel, channel_sum_kernel   the bias of the wide layers that stay on the library GEMM
//
// Common to the three tcgen05 kernels:
//   * forward orientation: UMMA M = 128 POINTS (TMEM lane = point), N = output channels (TMEM column), K = input
//     channels.  With the points on the lanes, the epilogue's stores are coalesced as they come out of tcgen05.ld (lane =
//     point, register = channel: a warp writes 128 contiguous bytes of one channel row per register), the bias is a
//     per-register scalar, and a thin layer (4 or 16 output channels) pads N to 16 instead of M to 128.
//   * operands in shared memory in the canonical K-major, no-swizzle layout (8-row x 16-byte core matrices):
//       byte offset of (row r, channel k) = (k / 4) * LBO + (r / 8) * 128 + (r % 8) * 16 + (k % 4) * 4,  LBO = rows * 16
//     A (points x channels) is the transpose of how x lies in HBM (channel rows, points contiguous): a thread reads
//     consecutive channels of ONE point, each load coalesced across the warp's 32 points, rounds to TF32 (cvt.rna, what
//     cuDNN / cuBLAS feed their TF32 kernels) and writes 16-byte core-matrix rows; the warp's 32 stores cover 512
//     contiguous bytes: no bank conflicts.  B (output channel x input channel) is w as it lies.
//   * steps of 32 input channels: tcgen05.mma is asynchronous; one thread issues 4 MMAs of K = 8 per step and commits
//     them to the stage's mbarrier, the other threads are already loading the next steps.
//   * accumulator: 128 lanes x N columns of fp32 in TMEM; epilogue warp w reads lanes 32 (w % 4)... (its quarter) and
//     every (warps / 4)-th block of 32 columns.
// Roofline: HBM for the thin layers (64 -> 256 channels over 64 x 3072 points: 50 MB in, 201 MB out), TF32 tensor
// throughput for the wide ones (512 -> 1024: 137 GFLOP) — those are left to the library (see pointwise_conv_kernel).
#include <cstdlib>

#include "sm100.cuh"

namespace mvp {

constexpr int kPwM = 128;       // points per CTA (UMMA M)
constexpr int kPwKC = 32;       // input channels per stage
constexpr int kPwThreads = 256;
constexpr int kPwMaxN = 256;    // output channels per CTA (UMMA N)

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// shared-memory matrix descriptor, no swizzle: start address, leading (K) and stride (8-row group) byte offsets in
// 16-byte units, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// four consecutive input channels of one output channel's weight row, zero beyond either extent
__device__ __forceinline__ float4 load_w4(const float *__restrict__ w, int cin, int cout, int co, int k, bool vec) {
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (co < cout) {
    const float *wr = w + (size_t)co * cin + k;
    if (vec && k + 3 < cin) {
      q = __ldg(reinterpret_cast<const float4 *>(wr));
    } else {
      if (k + 0 < cin) q.x = __ldg(wr + 0);
      if (k + 1 < cin) q.y = __ldg(wr + 1);
      if (k + 2 < cin) q.z = __ldg(wr + 2);
      if (k + 3 < cin) q.w = __ldg(wr + 3);
    }
  }
  return q;
}

// x (b, cin, n), w (cout, cin), bias (cout) or null, y (b, cout, n).  grid (ceil(n / 128), ceil(cout / nt), b).
// nt: output channels per CTA (multiple of 16, <= 256); tmem_cols: power of two >= max(32, nt).
// mask (b, cin, n) or null: x is multiplied by (mask > 0) as it is staged — the ReLU derivative of the backward pass.
__global__ void __launch_bounds__(kPwThreads) pointwise_conv_kernel(int cin, int cout, int n, const float *__restrict__ x,
                                                                    const float *__restrict__ mask,
                                                                    const float *__restrict__ w,
                                                                    const float *__restrict__ bias, float *__restrict__ y,
                                                                    int relu, int nt, int tmem_cols) {
  extern __shared__ __align__(128) unsigned char pw_smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, co0 = blockIdx.y * nt, p0 = blockIdx.x * kPwM;
  const uint32_t a_bytes = kPwM * kPwKC * 4, b_bytes = (uint32_t)nt * kPwKC * 4;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t smem0 = (smem_u32(pw_smem) + 127u) & ~127u;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  // instruction descriptor: D fp32, A and B TF32, both K-major, N = nt, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(kPwM >> 4) << 24);
  const float *xb = x + (size_t)b * cin * n;
  const float *mb = mask ? mask + (size_t)b * cin * n : nullptr;
  const int nchunk = (cin + kPwKC - 1) / kPwKC;
  const int pm = tid & (kPwM - 1);          // this thread's point of the tile (A staging)
  const bool p_ok = p0 + pm < n;
  const bool w_vec = (cin & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0;

  for (int c = 0; c < nchunk; c++) {
    const int s = c & 1;
    const int k0 = c * kPwKC;
    const int kc_n = (min(kPwKC, cin - k0) + 3) >> 2;  // 4-channel groups of this stage that hold anything
    const int kq_n = (kc_n + 1) >> 1;                  // MMAs (K = 8) of this stage
    if (c >= 2) mbar_wait(&bar[s], ((c >> 1) - 1) & 1);  // the MMAs that read this stage's previous contents are done
    const uint32_t a_s = smem0 + s * stage_bytes, b_s = a_s + a_bytes;
    // ---- A: points x channels.  (group g, point pm): g = (tid >> 7) + 2 i
    {
      float v[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int g = (tid >> 7) + 2 * i;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int k = k0 + g * 4 + j;
          float t = 0.f;
          if (g < 2 * kq_n && k < cin && p_ok) {
            t = __ldg(xb + (size_t)k * n + p0 + pm);
            if (mb && !(__ldg(mb + (size_t)k * n + p0 + pm) > 0.f)) t = 0.f;
          }
          v[i][j] = t;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int g = (tid >> 7) + 2 * i;
        if (g < 2 * kq_n)
          sts128(a_s + g * (kPwM * 16) + pm * 16, to_tf32(v[i][0]), to_tf32(v[i][1]), to_tf32(v[i][2]), to_tf32(v[i][3]));
      }
    }
    // ---- B: output channels x input channels.  item = g * nt + r; four items' loads in flight per thread
    for (int base = 0; base < 2 * kq_n * nt; base += 4 * kPwThreads) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int item = base + u * kPwThreads + tid;
        t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item < 2 * kq_n * nt) t[u] = load_w4(w, cin, cout, co0 + item % nt, k0 + (item / nt) * 4, w_vec);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int item = base + u * kPwThreads + tid;
        if (item < 2 * kq_n * nt)
          sts128(b_s + (item / nt) * (nt * 16) + (item % nt) * 16, to_tf32(t[u].x), to_tf32(t[u].y), to_tf32(t[u].z), to_tf32(t[u].w));
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int q = 0; q < kq_n; q++) {
        const uint64_t ad = umma_desc(a_s + q * 2 * (kPwM * 16), kPwM * 16, 128);
        const uint64_t bd = umma_desc(b_s + q * 2 * (nt * 16), nt * 16, 128);
        umma_tf32(tmem, ad, bd, idesc, (c > 0 || q > 0) ? 1u : 0u);
      }
      umma_commit(&bar[s]);  // arrives once every MMA issued so far has completed
    }
  }
  {
    const int last = nchunk - 1;
    mbar_wait(&bar[last & 1], (last >> 1) & 1);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: lane = point, register = output channel
  const int q = warp & 3;
  const int p = p0 + q * 32 + lane;
  float *yb = y + (size_t)b * cout * n;
  for (int cb = warp >> 2; cb * 32 < nt; cb += 2) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + cb * 32, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const int col = cb * 32 + j, co = co0 + col;
      if (col < nt && co < cout && p < n) {
        float v = __uint_as_float(r[j]);
        if (bias) v += __ldg(bias + co);
        if (relu) v = fmaxf(v, 0.f);
        yb[(size_t)co * n + p] = v;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

// ---- the thin layers: weight resident, persistent CTAs -------------------------------------------------------------------
// When the (nt x cin) weight tile fits in shared memory next to the staging ring it is staged ONCE per CTA, and the CTA
// walks over (cloud, 128-point tile) pairs in steps of 32 input channels.  A step's features travel HBM -> registers
// (eight coalesced 4-byte loads per thread: one point, eight channels) -> TF32 rounding -> two 16-byte stores into the
// K-major core-matrix layout; the loads of the next kPwAhead steps are already in flight in registers when a step is
// stored (48 KB per SM).  One thread issues the step's MMAs into one of TWO accumulators in tensor memory, and while
// the tensor core works on tile T all 16 warps run the epilogue of tile T - 1 from the other accumulator.
// (Tried and dropped: cp.async 4-byte copies straight into the layout — LDGSTS.32 moves about one element per clock per
// SM, 1.7 us per 16 KB step; 16-byte copies need the MN-major operand form, which in its no-swizzle variant gave zeros.)
constexpr int kPwStages = 3;   // shared-memory stages (the tensor core reads one while the next is written)
constexpr int kPwAhead = 3;    // steps of loads in flight in registers
constexpr int kPwRThreads = 512;
static_assert(kPwStages == kPwAhead, "stage u and register set u belong to the same step");

// grid (G, channel tiles): CTA (g, ct) owns output channels [ct * nt, ct * nt + nt) and the point tiles g, g + G, ...
// of the b * ceil(n / 128) tiles.  acc_cols: columns between the two accumulators (power of two >= max(32, nt)).
__global__ void __launch_bounds__(kPwRThreads, 1) pointwise_resident_kernel(int b, int cin, int cout, int n,
                                                                            const float *__restrict__ x,
                                                                            const float *__restrict__ w,
                                                                            const float *__restrict__ bias,
                                                                            float *__restrict__ y, int relu, int nt,
                                                                            int acc_cols) {
  extern __shared__ __align__(128) unsigned char pw_smem[];
  __shared__ __align__(8) uint64_t abar[kPwStages];  // a stage's MMAs are done: it may be overwritten
  __shared__ __align__(8) uint64_t tbar[2];          // an accumulator is complete
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[kPwMaxN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int co0 = blockIdx.y * nt;
  const int cin8 = (cin + 7) & ~7;
  const int nk = (cin + kPwKC - 1) / kPwKC;  // steps per tile
  const int tiles_p = (n + kPwM - 1) / kPwM;
  const int tiles = b * tiles_p;
  const int my_tiles = tiles > (int)blockIdx.x ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int steps = my_tiles * nk;
  const uint32_t a_bytes = kPwM * kPwKC * 4;
  const uint32_t smem0 = (smem_u32(pw_smem) + 127u) & ~127u;
  const uint32_t w_s = smem0 + kPwStages * a_bytes;  // [cin8 / 4][nt][16 bytes]

  if (tid == 0) {
    for (int i = 0; i < kPwStages; i++) mbar_init(&abar[i], 1);
    mbar_init(&tbar[0], 1);
    mbar_init(&tbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(2 * acc_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < nt; i += kPwRThreads) s_bias[i] = (bias && co0 + i < cout) ? __ldg(bias + co0 + i) : 0.f;

  // the features of a step (32 input channels of one 128-point tile): this thread's are point pm, channels kgrp * 8 ... + 7.
  // The load side runs kPwAhead steps ahead of the compute side; both keep their (tile, chunk) position incrementally
  // (the kernel is bound by instruction issue: no divisions or 64-bit multiplies per step).
  const int pm = tid & (kPwM - 1), kgrp = tid >> 7;
  const size_t chunk_stride = (size_t)kPwKC * n;
  int ld_c = 0, ld_tile = blockIdx.x;
  bool ld_pok = false;
  const float *ld_ptr = x;
  auto ld_setup = [&]() {  // at the start of a tile
    const int bb = ld_tile / tiles_p, p0 = (ld_tile - bb * tiles_p) * kPwM;
    ld_pok = p0 + pm < n;
    ld_ptr = x + ((size_t)bb * cin + kgrp * 8) * n + (ld_pok ? p0 + pm : 0);
  };
  ld_setup();
  auto gload = [&](float (&r)[8]) {
    const int kleft = cin - (ld_c * kPwKC + kgrp * 8);  // channels of this thread's eight that exist
    if (ld_pok && kleft >= 8) {
#pragma unroll
      for (int i = 0; i < 8; i++) r[i] = __ldg(ld_ptr + (size_t)i * n);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) r[i] = (ld_pok && i < kleft) ? __ldg(ld_ptr + (size_t)i * n) : 0.f;
    }
    ld_ptr += chunk_stride;
    if (++ld_c == nk) {
      ld_c = 0;
      ld_tile += gridDim.x;
      ld_setup();
    }
  };
  float rg[kPwAhead][8];
#pragma unroll
  for (int u = 0; u < kPwAhead; u++)
    if (u < steps) gload(rg[u]);

  // the weight tile, once: rows co0 ... co0 + nt - 1, all (padded) input channels
  {
    const bool w_vec = (cin & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0;
    const int items = (cin8 >> 2) * nt;
    for (int base = 0; base < items; base += 4 * kPwRThreads) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int item = base + u * kPwRThreads + tid;
        t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item < items) t[u] = load_w4(w, cin, cout, co0 + item % nt, (item / nt) * 4, w_vec);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int item = base + u * kPwRThreads + tid;
        if (item < items)
          sts128(w_s + (item / nt) * (nt * 16) + (item % nt) * 16, to_tf32(t[u].x), to_tf32(t[u].y), to_tf32(t[u].z), to_tf32(t[u].w));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(kPwM >> 4) << 24);

  // lane = point, register = output channel; warp: lane quarter warp % 4, column blocks warp / 4, warp / 4 + 4, ...
  auto epilogue = [&](int tl) {
    mbar_wait(&tbar[tl & 1], (tl >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int tile = blockIdx.x + tl * gridDim.x;
    const int bb = tile / tiles_p, p0 = (tile - bb * tiles_p) * kPwM, p = p0 + (warp & 3) * 32 + lane;
    float *yb = y + ((size_t)bb * cout + co0) * n + p;
    const bool full = p0 + kPwM <= n && co0 + nt <= cout;  // no guards needed anywhere in the tile
    for (int cb = warp >> 2; cb * 32 < nt; cb += kPwRThreads / 128) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (tl & 1) * acc_cols + cb * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float *yp = yb + (size_t)(cb * 32) * n;
      if (full && cb * 32 + 32 <= nt) {
        const float4 *sb = reinterpret_cast<const float4 *>(s_bias + cb * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++) {
          const float4 bv = sb[j4];
          float v0 = __uint_as_float(r[4 * j4 + 0]) + bv.x, v1 = __uint_as_float(r[4 * j4 + 1]) + bv.y;
          float v2 = __uint_as_float(r[4 * j4 + 2]) + bv.z, v3 = __uint_as_float(r[4 * j4 + 3]) + bv.w;
          if (relu) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f), v2 = fmaxf(v2, 0.f), v3 = fmaxf(v3, 0.f);
          yp[(size_t)(4 * j4 + 0) * n] = v0;
          yp[(size_t)(4 * j4 + 1) * n] = v1;
          yp[(size_t)(4 * j4 + 2) * n] = v2;
          yp[(size_t)(4 * j4 + 3) * n] = v3;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int col = cb * 32 + j;
          if (col < nt && co0 + col < cout && p < n) {
            float v = __uint_as_float(r[j]) + s_bias[col];
            if (relu) v = fmaxf(v, 0.f);
            yp[(size_t)j * n] = v;
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  int tl = 0, c = 0;  // the compute side's tile (of this CTA) and chunk
#pragma unroll 1
  for (int q0 = 0; q0 < steps; q0 += kPwAhead) {
#pragma unroll
    for (int u = 0; u < kPwAhead; u++) {  // kPwAhead == kPwStages: stage u, registers u
      const int q = q0 + u;
      if (q >= steps) break;  // uniform over the CTA
      // the stage's previous contents (step q - kPwStages) have been consumed by the tensor core
      if (q0 > 0) mbar_wait(&abar[u], ((q0 / kPwStages) - 1) & 1);
      const uint32_t a_s = smem0 + u * a_bytes;
      {
        const uint32_t mine = a_s + (kgrp * 2) * (kPwM * 16) + pm * 16;
        sts128(mine, to_tf32(rg[u][0]), to_tf32(rg[u][1]), to_tf32(rg[u][2]), to_tf32(rg[u][3]));
        sts128(mine + kPwM * 16, to_tf32(rg[u][4]), to_tf32(rg[u][5]), to_tf32(rg[u][6]), to_tf32(rg[u][7]));
      }
      if (q + kPwAhead < steps) gload(rg[u]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int kq_n = min(kPwKC, cin8 - c * kPwKC) >> 3;
        for (int j = 0; j < kq_n; j++) {
          const uint64_t ad = umma_desc(a_s + j * 2 * (kPwM * 16), kPwM * 16, 128);
          const uint64_t bd = umma_desc(w_s + (c * (kPwKC / 4) + j * 2) * (nt * 16), nt * 16, 128);
          umma_tf32(tmem + (tl & 1) * acc_cols, ad, bd, idesc, (c > 0 || j > 0) ? 1u : 0u);
        }
        umma_commit(&abar[u]);
        if (c == nk - 1) umma_commit(&tbar[tl & 1]);
      }
      if (++c == nk) {
        c = 0;
        if (tl > 0) epilogue(tl - 1);  // overlaps the tensor core's work on tile tl
        tl++;
      }
    }
  }
  if (my_tiles > 0) epilogue(my_tiles - 1);
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * acc_cols) : "memory");
}

static int pointwise_launch(int b, int cin, int cout, int n, const float *x, const float *mask, const float *w,
                            const float *bias, int relu, float *y, cudaStream_t s) {
  if (b < 0 || cin <= 0 || cout <= 0 || n < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0) return MVP_OK;
  if (b > 65535 || !x || !w || !y) return MVP_ERR_INVALID_ARGUMENT;
  const int tiles_n = (cout + kPwMaxN - 1) / kPwMaxN;
  int nt = (cout + tiles_n - 1) / tiles_n;  // even split over the channel tiles
  nt = (nt + 15) & ~15;
  int cols = 32;
  while (cols < nt) cols <<= 1;
  // the weight tile resident next to the staging ring (thin layers): persistent CTAs, two accumulators
  static const int mode = [] { const char *e = getenv("MVP_POINTWISE"); return e ? atoi(e) : 0; }();  // 1: streaming kernel only
  const size_t ring = (size_t)kPwStages * kPwM * kPwKC * 4, cin8 = (size_t)((cin + 7) & ~7), budget = 200 * 1024;
  auto fits = [&](int ntx) { return ring + cin8 * ntx * 4 + 128 <= budget; };
  int ct = tiles_n, ntr = nt;  // more, narrower channel tiles until the weight tile fits (x is then re-read from L2)
  while (!fits(ntr) && ntr > 32) {
    ct++;
    ntr = (((cout + ct - 1) / ct) + 15) & ~15;
  }
  if (!mask && mode != 1 && fits(ntr) && ct <= 8) {
    const size_t smem_r = ring + cin8 * ntr * 4 + 128;
    int colsr = 32;
    while (colsr < ntr) colsr <<= 1;
    static size_t granted_r[kMaxDevices];
    const int st = grant_dyn_smem(pointwise_resident_kernel, smem_r, granted_r);
    if (st != MVP_OK) return st;
    ct = (cout + ntr - 1) / ntr;
    const long long tiles = (long long)b * ((n + kPwM - 1) / kPwM);
    dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(tiles, kNumSMs / ct)), ct, 1);
    pointwise_resident_kernel<<<grid, kPwRThreads, smem_r, s>>>(b, cin, cout, n, x, w, bias, y, relu, ntr, colsr);
    count_launch();
    return launch_status();
  }
  const size_t smem = 2 * ((size_t)kPwM * kPwKC * 4 + (size_t)nt * kPwKC * 4) + 128;
  static size_t granted[kMaxDevices];
  const int st = grant_dyn_smem(pointwise_conv_kernel, smem, granted);
  if (st != MVP_OK) return st;
  dim3 grid((n + kPwM - 1) / kPwM, (cout + nt - 1) / nt, b);
  pointwise_conv_kernel<<<grid, kPwThreads, smem, s>>>(cin, cout, n, x, mask, w, bias, y, relu, nt, cols);
  count_launch();
  return launch_status();
}

// ---- the weight (and bias) gradient of a 1x1 layer --------------------------------------------------------------------------
// gw[o, c] = sum over clouds and points of g[b, o, p] * x[b, c, p]: UMMA M = output channels (128 per CTA row tile),
// N = input channels (<= 256), K = points — and both operands are K-major AS THEY LIE in HBM (points contiguous), so a
// 16-byte load of four points of one channel row is one core-matrix row.  With a bias the B operand gets one more row
// of ones: column cin of the product is sum_p g[b, o, p], the bias gradient, for free.
// Persistent CTAs split the b * ceil(n / 32) K-steps evenly; each accumulates its share in tensor memory and writes one
// (128 x N) partial; pointwise_wgrad_reduce_kernel adds the partials in a fixed order (deterministic).
// Loads run two steps ahead in registers (up to 6 x 16 bytes per thread and step), rounded to TF32 on the way.
// Measured alternatives (clock64 per phase, 64 x (256 x 64) x 3072): a cp.async ring of 16-byte copies, five steps in
// flight, was SLOWER (0.156 against 0.116 ms): issuing a step's ~1700 copies takes ~1400 cycles (LDGSTS retires about
// one request per clock per SM) and the in-place rounding pass + proxy fence another ~740, all of it serialised with
// the other phases because every warp of the CTA runs the same phase at the same time.  What comes next is a
// producer / consumer split of the warps (or TMA for the copies), not more copies in flight.
constexpr int kWgThreads = 512;
constexpr int kWgStages = 3;
constexpr int kWgKC = 32;  // points per step

template <bool kVec>
__device__ __forceinline__ float4 load_p4(const float *__restrict__ row, int p, int n) {
  if (kVec) return p < n ? __ldg(reinterpret_cast<const float4 *>(row + p)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p + 0 < n) q.x = __ldg(row + p + 0);
  if (p + 1 < n) q.y = __ldg(row + p + 1);
  if (p + 2 < n) q.z = __ldg(row + p + 2);
  if (p + 3 < n) q.w = __ldg(row + p + 3);
  return q;
}

// grid (G, row tiles of 128 output channels).  nb: rows of the B operand = UMMA N (multiple of 16 >= cin + ones);
// ones: 1 if row cin of B is the row of ones.  partial: [gridDim.y][G][128][nb].
// kVec: n % 4 == 0 and 16-byte aligned g, x (else scalar loads, same layout).
template <bool kVec>
__global__ void __launch_bounds__(kWgThreads, 1) pointwise_wgrad_kernel(int b, int cin, int cout, int n,
                                                                        const float *__restrict__ g,
                                                                        const float *__restrict__ x, int nb, int ones,
                                                                        int tmem_cols, float *__restrict__ partial) {
  extern __shared__ __align__(128) unsigned char pw_smem[];
  __shared__ __align__(8) uint64_t abar[kWgStages];
  __shared__ __align__(8) uint64_t dbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int o0 = blockIdx.y * 128;
  const int rows_a = min(128, cout - o0);
  const int steps_c = (n + kWgKC - 1) / kWgKC;               // K-steps per cloud
  const long long total = (long long)b * steps_c;
  const long long s_lo = total * blockIdx.x / gridDim.x, s_hi = total * (blockIdx.x + 1) / gridDim.x;
  const int steps = (int)(s_hi - s_lo);
  const uint32_t a_bytes = 128 * kWgKC * 4, b_bytes = (uint32_t)nb * kWgKC * 4, st_bytes = a_bytes + b_bytes;
  const uint32_t smem0 = (smem_u32(pw_smem) + 127u) & ~127u;

  if (tid == 0) {
    for (int i = 0; i < kWgStages; i++) mbar_init(&abar[i], 1);
    mbar_init(&dbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows that no store ever writes: zero (A rows >= rows_a, B rows >= cin), or one (B row cin with a bias)
  for (uint32_t i = tid; i < kWgStages * st_bytes / 16; i += kWgThreads) {
    const uint32_t off = i * 16, in_st = off % st_bytes;
    float v = 0.f;
    if (in_st >= a_bytes) {
      const int r = ((in_st - a_bytes) % (nb * 16)) / 16;
      if (ones && r == cin) v = 1.f;
    }
    const uint32_t u = __float_as_uint(v);
    sts128(smem0 + off, u, u, u, u);
  }

  // Fixed per thread for the whole kernel (nothing is recomputed per step): for each of its chunks (row r, 4-point
  // group kc) the row's offset inside a cloud (-1: nothing to load) and the byte offset inside a stage with kc in the
  // low bits.  A warp covers 8 rows x 4 groups (64 contiguous bytes per row in HBM, 128 contiguous bytes per 8 rows in
  // shared memory): item = tid + 512 i, rl = item % 8, kc4 = (item / 8) % 4, rh = item / 32,
  // row = rh % (rows / 8) * 8 + rl, kc = kc4 + 4 * (rh / (rows / 8)).
  int a_row[2], a_so[2], b_row[4], b_so[4];
  {
    auto rowkc = [&](int item, int rows, int &r, int &kc) {
      const int rl = item & 7, kc4 = (item >> 3) & 3, rh = item >> 5, rg = rows >> 3;
      r = (rh % rg) * 8 + rl;
      kc = kc4 + 4 * (rh / rg);
    };
#pragma unroll
    for (int i = 0; i < 2; i++) {
      int r, kc;
      rowkc(tid + i * kWgThreads, 128, r, kc);
      a_row[i] = r < rows_a ? r * n : -1;
      a_so[i] = (kc * (128 * 16) + r * 16) | kc;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int r, kc;
      rowkc(tid + i * kWgThreads, nb, r, kc);
      b_row[i] = (kc < 8 && r < cin) ? r * n : -1;
      b_so[i] = ((int)a_bytes + (kc & 7) * (nb * 16) + r * 16) | (kc & 7);
    }
  }
  int ld_p0;  // the load side's first point of the step, in its cloud
  const float *ld_g, *ld_x;
  {
    const int ld_b = (int)(s_lo / steps_c);
    ld_p0 = (int)(s_lo - (long long)ld_b * steps_c) * kWgKC;
    ld_g = g + ((size_t)ld_b * cout + o0) * n;
    ld_x = x + (size_t)ld_b * cin * n;
  }
  auto gload = [&](float4 (&ra)[2], float4 (&rb)[4]) {
#pragma unroll
    for (int i = 0; i < 2; i++)
      ra[i] = a_row[i] >= 0 ? load_p4<kVec>(ld_g + a_row[i], ld_p0 + (a_so[i] & 7) * 4, n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; i++)
      rb[i] = b_row[i] >= 0 ? load_p4<kVec>(ld_x + b_row[i], ld_p0 + (b_so[i] & 7) * 4, n) : make_float4(0.f, 0.f, 0.f, 0.f);
    ld_p0 += kWgKC;
    if (ld_p0 >= n) {
      ld_p0 = 0;
      ld_g += (size_t)cout * n;
      ld_x += (size_t)cin * n;
    }
  };
  float4 ra[2][2], rb[2][4];
#pragma unroll
  for (int u = 0; u < 2; u++)
    if (u < steps) gload(ra[u], rb[u]);

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nb >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  auto step = [&](int q, int sg, float4 (&qa)[2], float4 (&qb)[4]) {
    if (q >= kWgStages) mbar_wait(&abar[sg], ((q / kWgStages) - 1) & 1);
    const uint32_t a_s = smem0 + sg * st_bytes, b_s = a_s + a_bytes;
#pragma unroll
    for (int i = 0; i < 2; i++)
      if (a_row[i] >= 0)
        sts128(a_s + (a_so[i] & ~15), to_tf32(qa[i].x), to_tf32(qa[i].y), to_tf32(qa[i].z), to_tf32(qa[i].w));
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (b_row[i] >= 0)
        sts128(a_s + (b_so[i] & ~15), to_tf32(qb[i].x), to_tf32(qb[i].y), to_tf32(qb[i].z), to_tf32(qb[i].w));
    if (q + 2 < steps) gload(qa, qb);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int j = 0; j < kWgKC / 8; j++) {
        const uint64_t ad = umma_desc(a_s + j * 2 * (128 * 16), 128 * 16, 128);
        const uint64_t bd = umma_desc(b_s + j * 2 * (nb * 16), nb * 16, 128);
        umma_tf32(tmem, ad, bd, idesc, (q > 0 || j > 0) ? 1u : 0u);
      }
      umma_commit(&abar[sg]);
      if (q == steps - 1) umma_commit(&dbar);
    }
  };
  // stages cycle with period 3, register sets with period 2: unroll by 6
#pragma unroll 1
  for (int q0 = 0; q0 < steps; q0 += 6) {
#pragma unroll
    for (int u = 0; u < 6; u++) {
      if (q0 + u >= steps) break;
      step(q0 + u, u % 3, ra[u & 1], rb[u & 1]);
    }
  }
  // ---- the partial: lane = output channel, register = input channel
  float *part = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 128 * nb;
  if (steps > 0) {
    mbar_wait(&dbar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const int row = (warp & 3) * 32 + lane;
  for (int cb = warp >> 2; cb * 32 < nb; cb += kWgThreads / 128) {
    uint32_t r[32];
    if (steps > 0) {
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + cb * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int j = 0; j < 32; j++) r[j] = 0u;
    }
    float4 *dst = reinterpret_cast<float4 *>(part + (size_t)row * nb + cb * 32);
#pragma unroll
    for (int j4 = 0; j4 < 8; j4++)
      if (cb * 32 + j4 * 4 < nb)
        dst[j4] = make_float4(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1]), __uint_as_float(r[4 * j4 + 2]),
                              __uint_as_float(r[4 * j4 + 3]));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

// gw[o, c] (and gb[o] = column cin) = sum over the G partials, in order
__global__ void pointwise_wgrad_reduce_kernel(int cin, int cout, int nb, int G, int ones, const float *__restrict__ partial,
                                              float *__restrict__ gw, float *__restrict__ gb) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, o = blockIdx.y;
  if (c >= cin + ones) return;
  const float *p = partial + ((size_t)(o >> 7) * G * 128 + (o & 127)) * nb + c;
  float t = 0.f;
  const size_t stride = (size_t)128 * nb;
  int i = 0;
  for (; i + 8 <= G; i += 8) {  // eight loads in flight, added in order
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = __ldg(p + (size_t)(i + u) * stride);
#pragma unroll
    for (int u = 0; u < 8; u++) t += v[u];
  }
  for (; i < G; i++) t += __ldg(p + (size_t)i * stride);
  if (c < cin) gw[(size_t)o * cin + c] = t;
  else if (gb) gb[o] = t;
}

// ---- the bias of the wide layers -------------------------------------------------------------------------------------------
// The wide layers (512 -> 1024 channels over 64 x 2048 points: 137 GFLOP) stay on the library's TF32 GEMM; what this
// file contributes there is their bias, which torch adds with its generic broadcasting kernel at 2.7 TB/s and sums (for
// the gradient) at 2.1 TB/s: one row per CTA, float4 accesses, nothing else.
__global__ void __launch_bounds__(128) bias_add_kernel(int c, int n, float *__restrict__ y, const float *__restrict__ bias, int relu) {
  const size_t row = blockIdx.x;
  const float bv = __ldg(bias + (int)(row % c));
  float *yr = y + row * n;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    float4 *y4 = reinterpret_cast<float4 *>(yr);
    for (int i = threadIdx.x; i < (n >> 2); i += 128) {
      float4 v = y4[i];
      v.x += bv, v.y += bv, v.z += bv, v.w += bv;
      if (relu) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
      y4[i] = v;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += 128) {
      float v = yr[i] + bv;
      yr[i] = relu ? fmaxf(v, 0.f) : v;
    }
  }
}

// partial[ch * splits + s] = sum over clouds s, s + splits, ... and all points of g[., ch, .] (fixed order: deterministic)
__global__ void __launch_bounds__(256) channel_sum_kernel(int b, int c, int n, const float *__restrict__ g, float *__restrict__ partial) {
  const int ch = blockIdx.x, s = blockIdx.y, splits = gridDim.y;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bool vec = (n & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  for (int bb = s; bb < b; bb += splits) {
    const float *gr = g + ((size_t)bb * c + ch) * n;
    if (vec) {
      const float4 *g4 = reinterpret_cast<const float4 *>(gr);
      for (int i = threadIdx.x; i < (n >> 2); i += 256) {
        const float4 v = __ldg(g4 + i);
        acc[0] += v.x, acc[1] += v.y, acc[2] += v.z, acc[3] += v.w;
      }
    } else {
      for (int i = threadIdx.x; i < n; i += 256) acc[0] += __ldg(gr + i);
    }
  }
  float v = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ float sw[8];
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += sw[i];
    partial[(size_t)ch * splits + s] = t;
  }
}
__global__ void channel_sum_final_kernel(int c, int splits, const float *__restrict__ partial, float *__restrict__ out) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float t = 0.f;
  for (int s = 0; s < splits; s++) t += partial[(size_t)ch * splits + s];
  out[ch] = t;
}

}  // namespace mvp

MVP_API int mvp_pointwise_conv(int b, int cin, int cout, int n, const float *x, const float *w, const float *bias, int relu,
                               float *y, mvp_stream_t stream) {
  return mvp::pointwise_launch(b, cin, cout, n, x, nullptr, w, bias, relu, y, (cudaStream_t)stream);
}

MVP_API int mvp_pointwise_conv_masked(int b, int cin, int cout, int n, const float *x, const float *mask, const float *w,
                                      float *y, mvp_stream_t stream) {
  if (!mask && b > 0 && n > 0) return MVP_ERR_INVALID_ARGUMENT;
  return mvp::pointwise_launch(b, cin, cout, n, x, mask, w, nullptr, 0, y, (cudaStream_t)stream);
}

MVP_API int mvp_bias_add(int b, int c, int n, float *y, const float *bias, int relu, mvp_stream_t stream) {
  if (b < 0 || c <= 0 || n < 0) return MVP_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0) return MVP_OK;
  if (!y || !bias || (unsigned long long)b * c > 0x7fffffffULL) return MVP_ERR_INVALID_ARGUMENT;
  mvp::bias_add_kernel<<<(unsigned)((size_t)b * c), 128, 0, (cudaStream_t)stream>>>(c, n, y, bias, relu);
  mvp::count_launch();
  return mvp::launch_status();
}

MVP_API size_t mvp_channel_sum_workspace_bytes(int b, int c) {
  if (b <= 0 || c <= 0) return 0;
  const int splits = std::max(1, std::min(b, (4 * mvp::kNumSMs + c - 1) / c));
  return (size_t)c * splits * sizeof(float);
}

MVP_API int mvp_channel_sum(int b, int c, int n, const float *g, float *out, void *workspace, size_t workspace_bytes,
                            mvp_stream_t stream) {
  if (b <= 0 || c <= 0 || n <= 0 || !g || !out) return MVP_ERR_INVALID_ARGUMENT;
  if (!workspace || workspace_bytes < mvp_channel_sum_workspace_bytes(b, c)) return MVP_ERR_WORKSPACE;
  const int splits = std::max(1, std::min(b, (4 * mvp::kNumSMs + c - 1) / c));
  cudaStream_t s = (cudaStream_t)stream;
  mvp::channel_sum_kernel<<<dim3(c, splits), 256, 0, s>>>(b, c, n, g, (float *)workspace);
  mvp::channel_sum_final_kernel<<<(c + 127) / 128, 128, 0, s>>>(c, splits, (const float *)workspace, out);
  mvp::count_launch(2);
  return mvp::launch_status();
}

// the weight gradient's launch geometry: UMMA N, persistent CTAs per row tile, row tiles
static void wgrad_geometry(int cin, int cout, int with_bias, int *nb, int *G, int *rt) {
  *nb = (cin + (with_bias ? 1 : 0) + 15) & ~15;
  *rt = (cout + 127) / 128;
  *G = std::max(1, mvp::kNumSMs / *rt);
}

MVP_API size_t mvp_pointwise_wgrad_workspace_bytes(int cin, int cout, int with_bias) {
  if (cin <= 0 || cout <= 0 || cin + (with_bias ? 1 : 0) > 256) return 0;
  int nb, G, rt;
  wgrad_geometry(cin, cout, with_bias, &nb, &G, &rt);
  return (size_t)rt * G * 128 * nb * sizeof(float);
}

MVP_API int mvp_pointwise_wgrad(int b, int cin, int cout, int n, const float *g, const float *x, float *grad_w, float *grad_bias,
                                void *workspace, size_t workspace_bytes, mvp_stream_t stream) {
  const int with_bias = grad_bias != nullptr;
  if (b <= 0 || cin <= 0 || cout <= 0 || n <= 0 || cin + with_bias > 256 || !g || !x || !grad_w) return MVP_ERR_INVALID_ARGUMENT;
  if (!workspace || workspace_bytes < mvp_pointwise_wgrad_workspace_bytes(cin, cout, with_bias)) return MVP_ERR_WORKSPACE;
  int nb, G, rt;
  wgrad_geometry(cin, cout, with_bias, &nb, &G, &rt);
  int cols = 32;
  while (cols < nb) cols <<= 1;
  const size_t smem = (size_t)mvp::kWgStages * (128 + nb) * mvp::kWgKC * 4 + 128;
  const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x)) & 15) == 0;
  static size_t granted[2][mvp::kMaxDevices];
  auto kernel = vec ? mvp::pointwise_wgrad_kernel<true> : mvp::pointwise_wgrad_kernel<false>;
  const int st = mvp::grant_dyn_smem(kernel, smem, granted[vec]);
  if (st != MVP_OK) return st;
  cudaStream_t s = (cudaStream_t)stream;
  kernel<<<dim3(G, rt), mvp::kWgThreads, smem, s>>>(b, cin, cout, n, g, x, nb, with_bias, cols, (float *)workspace);
  mvp::pointwise_wgrad_reduce_kernel<<<dim3((cin + with_bias + 127) / 128, cout), 128, 0, s>>>(cin, cout, nb, G, with_bias,
                                                                                               (const float *)workspace, grad_w,
                                                                                               grad_bias);
  mvp::count_launch(2);
  return mvp::launch_status();
}
