// Chamfer distance, large-cloud path for sm_100a: BOTH directions from ONE evaluation of each pair.
//
// Replaces the two NmDistanceKernel launches of the reference (utils/metrics/CD/chamfer3D/chamfer3D.cu:12-134,
// 142-143), which evaluate every pair twice (once per direction).  The squared distance
//     d(i,j) = fma(dz,dz, fma(dx,dx, dy*dy)),  dx = xyz2[j].x - xyz1[i].x ...
// is bit-identical for the swapped direction (only the signs of dx,dy,dz flip, chamfer3D.cu:32-35), so one
// evaluation serves dist1/idx1 (row minima) and dist2/idx2 (column minima).
//
// Pass 1  chamfer_pair_kernel — a CTA owns a tile of TM = kBlk*WARPS rows (points of xyz1, kRows = 4 per thread, held in
//   registers as duplicated fp32x2 pairs) and sweeps column tiles of 1024 points of xyz2 staged in shared memory
//   (1-D bulk TMA of the raw xyz rows into a double buffer, then re-laid as x/y/z quads so that four columns are
//   three broadcast LDS.128).  Distances are computed two columns at a time with packed FADD2/FMUL2/FFMA2
//   (3 issue slots per pair instead of 6); minima are tracked as VALUES only, with three-input FMNMX3:
//     rows:    running minimum per row; per kBlk-column chunk (kBlk = 32*kRows = 128) a compare records WHICH
//              CHUNK improved it;
//     columns: per-thread minimum over its kRows rows, one redux.sync.min.u32 per column across the warp
//              (d >= +0, so the fp32 bit pattern orders like an unsigned integer), parked per warp in shared
//              memory, merged over the CTA's warps at the end of the tile.
//   Partial results meet in global memory as 64-bit keys  (bits(d) << 32) | block  through atomicMin: `block` is
//   the index of the kBlk-wide chunk of columns (for a row) or of rows (for a column) that produced the minimum,
//   so equal distances resolve to the LOWEST block — the reference's lowest-index tie rule at block granularity.
// Pass 2  chamfer_resolve_kernel — one warp per point re-evaluates the kBlk candidates of its winning block and
//   takes the first one whose distance has exactly the winning bit pattern: the lowest index among equal minima,
//   as the reference's strict `<` scan gives (chamfer3D.cu:36,126).  Costs kBlk/N of pass 1.
#include "common.cuh"
#include "sm100.cuh"

namespace mvp {

// Tuning knobs (tools/pair_variants.py builds the alternatives; the defaults are the measured best).
#ifndef MVP_PAIR_ROWS
#define MVP_PAIR_ROWS 4  // rows per thread (4: 72 registers, 3 CTAs/SM — measured best; 8: 126 registers, 2 CTAs/SM)
#endif
#ifndef MVP_PAIR_COLS
#define MVP_PAIR_COLS 4  // columns per inner step (2 or 4)
#endif
#ifndef MVP_PAIR_MINB
#define MVP_PAIR_MINB 3  // resident CTAs per SM asked of ptxas for the 8-warp tile
#endif
#ifndef MVP_PAIR_SCALAR
#define MVP_PAIR_SCALAR 0  // 1: scalar FADD/FMUL/FFMA instead of the packed fp32x2 forms
#endif

constexpr int kRows = MVP_PAIR_ROWS;
constexpr int kBlk = 32 * kRows;  // granularity of the "which block won" half of a key (rows and columns alike)
constexpr int kCps = MVP_PAIR_COLS;
constexpr int kTN = 1024;         // columns per shared-memory tile
constexpr float kPadRow = 2e19f, kPadCol = -2e19f;  // padding coordinates: any distance to them overflows to +inf

// two columns (packed) against one row (duplicated): the reference's contraction, lane by lane
__device__ __forceinline__ u64 dist2(u64 X, u64 Y, u64 Z, u64 qx, u64 qy, u64 qz) {
#if MVP_PAIR_SCALAR
  float xl, xh, yl, yh, zl, zh, ql, qh, rl, rh, sl, sh;
  unpack2(X, xl, xh), unpack2(Y, yl, yh), unpack2(Z, zl, zh);
  unpack2(qx, ql, qh), unpack2(qy, rl, rh), unpack2(qz, sl, sh);
  return pack2(sqdist(xl - ql, yl - rl, zl - sl), sqdist(xh - qh, yh - rh, zh - sh));
#else
  const u64 dx = sub2(X, qx), dy = sub2(Y, qy), dz = sub2(Z, qz);
  return fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
#endif
}

template <int WARPS>
struct PairSmem {
  float raw[2][kTN * 3];             // TMA landing zone: xyz rows as they lie in global memory (2 x 12 KB)
  ulonglong2 x[kTN / 4], y[kTN / 4], z[kTN / 4];  // the tile as quads: (x0,x1 | x2,x3) ...          (12 KB)
  uint32_t wmin[WARPS][kTN];           // per-warp column minima of the current tile                   (32 KB)
  uint64_t bar[2];
};

// A CTA's work item is (row tile, column split, cloud) = blockIdx: WARPS warps, TM = kBlk * WARPS rows per tile.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS >= 8 ? MVP_PAIR_MINB : (WARPS == 4 ? 2 * MVP_PAIR_MINB : 4))
chamfer_pair_kernel(int n, int m, int tiles_x, int split, int tiles_per_cta, int use_tma,
                    const float *__restrict__ xyz1, const float *__restrict__ xyz2, u64 *__restrict__ rowkey,
                    u64 *__restrict__ colkey) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PairSmem<WARPS> &S = *reinterpret_cast<PairSmem<WARPS> *>(smem_raw);
  constexpr int T = WARPS * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float inf = __int_as_float(0x7f800000);
  const int per_cloud = tiles_x * split;
  const long long work = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (use_tma && tid == 0) {
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t tt = 0;  // tiles fetched by TMA so far in this CTA: buffer = tt & 1, mbarrier parity = (tt >> 1) & 1

  {
  const int slot = (int)(work / per_cloud), rem = (int)(work % per_cloud);
  const int b = slot;
  const int bx = rem % tiles_x, by = rem / tiles_x;
  const float *A = xyz1 + (size_t)b * n * 3;
  const float *Bc = xyz2 + (size_t)b * m * 3;
  const int row0 = bx * (WARPS * kBlk) + warp * kBlk;  // first row of this warp's kBlk-row block
  const int tile0 = by * tiles_per_cta;
  const int ntiles_all = (m + kTN - 1) / kTN;
  const int ntiles = min(tiles_per_cta, ntiles_all - tile0);

  // ---- this thread's rows, duplicated into fp32x2 operands
  u64 qx[kRows], qy[kRows], qz[kRows];
  float best[kRows], prev[kRows];
  int bchunk[kRows];
#pragma unroll
  for (int r = 0; r < kRows; r++) {
    const int i = row0 + lane + 32 * r;
    float x = kPadRow, y = kPadRow, z = kPadRow;
    if (i < n) {
      x = __ldg(A + (size_t)i * 3 + 0);
      y = __ldg(A + (size_t)i * 3 + 1);
      z = __ldg(A + (size_t)i * 3 + 2);
    }
    qx[r] = pack2(x, x);
    qy[r] = pack2(y, y);
    qz[r] = pack2(z, z);
    best[r] = prev[r] = inf;
    bchunk[r] = 0;
  }

  // thread 0: fetch column tile t (a FULL tile) into raw[seq & 1], seq = its TMA sequence number in this CTA
  auto issue_tma = [&](int t, uint32_t seq) {
    const int buf = seq & 1;
    mbar_expect_tx(&S.bar[buf], kTN * 12);
    tma_load_1d(S.raw[buf], Bc + (size_t)(tile0 + t) * kTN * 3, kTN * 12, &S.bar[buf]);
  };
  // A tile goes through TMA when it is full and 16-byte aligned in global memory; else plain loads.
  auto tile_by_tma = [&](int t) { return use_tma && (tile0 + t + 1) * kTN <= m; };
  if (ntiles > 0 && tile_by_tma(0) && tid == 0) issue_tma(0, tt);

  for (int t = 0; t < ntiles; t++) {
    const int col0 = (tile0 + t) * kTN;
    // ---- stage the tile as x/y/z quads
    float *sx = reinterpret_cast<float *>(S.x), *sy = reinterpret_cast<float *>(S.y),
          *sz = reinterpret_cast<float *>(S.z);
    if (tile_by_tma(t)) {
      mbar_wait(&S.bar[tt & 1], (tt >> 1) & 1);
      const float *raw = S.raw[tt & 1];
      tt++;
      for (int j = tid; j < kTN; j += T) {
        sx[j] = raw[j * 3 + 0];
        sy[j] = raw[j * 3 + 1];
        sz[j] = raw[j * 3 + 2];
      }
    } else {
      for (int j = tid; j < kTN; j += T) {
        const int c = col0 + j;
        float x = kPadCol, y = kPadCol, z = kPadCol;
        if (c < m) {
          x = __ldg(Bc + (size_t)c * 3 + 0);
          y = __ldg(Bc + (size_t)c * 3 + 1);
          z = __ldg(Bc + (size_t)c * 3 + 2);
        }
        sx[j] = x;
        sy[j] = y;
        sz[j] = z;
      }
    }
    __syncthreads();
    if (t + 1 < ntiles && tile_by_tma(t + 1) && tid == 0) issue_tma(t + 1, tt);  // overlaps the sweep below

    // ---- sweep: kTN/kBlk chunks of kBlk columns, kCps columns per step
    const int cols_here = min(kTN, m - col0);
    const int nchunks = (cols_here + kBlk - 1) / kBlk;
    for (int c = 0; c < nchunks; c++) {
#pragma unroll 1
      for (int g = c * (kBlk / kCps); g < (c + 1) * (kBlk / kCps); g++) {
        u64 X[kCps / 2], Y[kCps / 2], Z[kCps / 2];
        if (kCps == 4) {
          const ulonglong2 x4 = S.x[g], y4 = S.y[g], z4 = S.z[g];
          X[0] = x4.x, X[kCps / 2 - 1] = x4.y, Y[0] = y4.x, Y[kCps / 2 - 1] = y4.y, Z[0] = z4.x, Z[kCps / 2 - 1] = z4.y;
        } else {
          X[0] = reinterpret_cast<const u64 *>(S.x)[g];
          Y[0] = reinterpret_cast<const u64 *>(S.y)[g];
          Z[0] = reinterpret_cast<const u64 *>(S.z)[g];
        }
        float cm[kCps];
#pragma unroll
        for (int h = 0; h < kCps; h++) cm[h] = inf;
#pragma unroll
        for (int r = 0; r < kRows; r += 2) {
#pragma unroll
          for (int h = 0; h < kCps / 2; h++) {
            const u64 d0 = dist2(X[h], Y[h], Z[h], qx[r], qy[r], qz[r]);
            const u64 d1 = dist2(X[h], Y[h], Z[h], qx[r + 1], qy[r + 1], qz[r + 1]);
            float d0l, d0h, d1l, d1h;
            unpack2(d0, d0l, d0h);
            unpack2(d1, d1l, d1h);
            best[r] = min3(best[r], d0l, d0h);
            best[r + 1] = min3(best[r + 1], d1l, d1h);
            cm[2 * h] = min3(cm[2 * h], d0l, d1l);
            cm[2 * h + 1] = min3(cm[2 * h + 1], d0h, d1h);
          }
        }
        uint32_t w[kCps];
#pragma unroll
        for (int h = 0; h < kCps; h++) w[h] = redux_min_u32(__float_as_uint(cm[h]));
        if (lane == 0) {
          if (kCps == 4)
            *reinterpret_cast<uint4 *>(&S.wmin[warp][g * 4]) = make_uint4(w[0], w[1], w[kCps - 2], w[kCps - 1]);
          else
            *reinterpret_cast<uint2 *>(&S.wmin[warp][g * 2]) = make_uint2(w[0], w[1]);
        }
      }
      const int chunk = col0 / kBlk + c;
#pragma unroll
      for (int r = 0; r < kRows; r++) {
        if (best[r] < prev[r]) {
          prev[r] = best[r];
          bchunk[r] = chunk;
        }
      }
    }
    __syncthreads();
    // ---- column keys of this tile: minimum over the CTA's warps, lowest row block on ties
    for (int j = tid; j < cols_here; j += T) {
      u64 k = ~0ull;
#pragma unroll
      for (int w = 0; w < WARPS; w++) {
        const u64 cand = ((u64)S.wmin[w][j] << 32) | (unsigned)((bx * WARPS + w));
        k = cand < k ? cand : k;
      }
      atomicMin(colkey + (size_t)b * m + col0 + j, k);
    }
    // (the next iteration's staging writes S.x/y/z, which nobody reads any more; S.wmin is rewritten only
    //  after the __syncthreads that follows the staging)
  }
  // ---- row keys
#pragma unroll
  for (int r = 0; r < kRows; r++) {
    const int i = row0 + lane + 32 * r;
    if (i < n && ntiles > 0)
      atomicMin(rowkey + (size_t)b * n + i, ((u64)__float_as_uint(best[r]) << 32) | (unsigned)bchunk[r]);
  }
  }
}

// One warp per point p (both directions in one launch: the b*n points of xyz1 against xyz2, then the b*m points of
// xyz2 against xyz1): key = (bits(d) << 32) | block.  The kBlk points of the other cloud in that block are
// re-evaluated, kBlk/32 consecutive candidates per lane (128-bit loads when the block is whole and 16-byte
// aligned), and the first one whose distance to p has exactly those bits is the answer.
//   grid: x strides over the n + m points of a cloud pair (8 per CTA), y over the clouds.
__global__ void __launch_bounds__(256)
chamfer_resolve_kernel(int b, int n, int m, int vec1, int vec2, const float *__restrict__ xyz1,
                       const float *__restrict__ xyz2, const u64 *__restrict__ rowkey, const u64 *__restrict__ colkey,
                       float *__restrict__ dist1, float *__restrict__ dist2, int *__restrict__ idx1,
                       int *__restrict__ idx2) {
  constexpr int PER = kBlk / 32;  // candidates per lane
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int slot = blockIdx.y; slot < b; slot += gridDim.y) {
  const long long cloud = slot;
  for (int i = blockIdx.x * 8 + warp; i < n + m; i += gridDim.x * 8) {
    const bool first_dir = i < n;
    const int np = first_dir ? n : m, nq = first_dir ? m : n;
    const int vec_ok = first_dir ? vec2 : vec1;  // alignment of the cloud the candidates come from
    const float *P = first_dir ? xyz1 : xyz2, *Q = first_dir ? xyz2 : xyz1;
    const long long wid = cloud * np + (first_dir ? i : i - n);
    const u64 k = __ldg((first_dir ? rowkey : colkey) + wid);
    const uint32_t bits = (uint32_t)(k >> 32);
    const int base = (int)(uint32_t)k * kBlk;
    const float px = __ldg(P + wid * 3 + 0), py = __ldg(P + wid * 3 + 1), pz = __ldg(P + wid * 3 + 2);
    const float *q = Q + ((size_t)cloud * nq + base + lane * PER) * 3;
    float c[PER * 3];
    if (vec_ok && base + kBlk <= nq) {
      const float4 *q4 = reinterpret_cast<const float4 *>(q);
#pragma unroll
      for (int v = 0; v < PER * 3 / 4; v++) {
        const float4 t = __ldg(q4 + v);
        c[v * 4 + 0] = t.x;
        c[v * 4 + 1] = t.y;
        c[v * 4 + 2] = t.z;
        c[v * 4 + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int v = 0; v < PER; v++) {
        const bool ok = base + lane * PER + v < nq;
        c[v * 3 + 0] = ok ? __ldg(q + v * 3 + 0) : kPadCol;
        c[v * 3 + 1] = ok ? __ldg(q + v * 3 + 1) : kPadCol;
        c[v * 3 + 2] = ok ? __ldg(q + v * 3 + 2) : kPadCol;
      }
    }
    int first = PER;
#pragma unroll
    for (int v = PER - 1; v >= 0; v--) {
      const float d = sqdist(c[v * 3 + 0] - px, c[v * 3 + 1] - py, c[v * 3 + 2] - pz);
      if (__float_as_uint(d) == bits) first = v;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, first < PER);
    const int src = mask ? __ffs(mask) - 1 : 0;
    const int v = __shfl_sync(0xffffffffu, first, src);
    if (lane == 0) {
      (first_dir ? dist1 : dist2)[wid] = __uint_as_float(bits);
      (first_dir ? idx1 : idx2)[wid] = mask ? base + src * PER + v : base;
    }
  }
  }
}

bool chamfer_fused_supported(int b, int n, int m) {
  // worth it once a cloud pair has enough work to amortise the key round-trip; tiny clouds use chamfer.cu
  return b > 0 && n >= 512 && m >= 512 && (long long)b <= 65535;
}

size_t chamfer_fused_workspace_bytes(int b, int n, int m) { return sizeof(u64) * (size_t)b * ((size_t)n + m); }

struct PairShape {
  int warps, tiles_per_cta, split, tiles_x;
};

// Tile shape: the largest row tile that still yields >= 4 CTAs per SM worth of tiles, then split the column sweep so
// that the tail of the last wave is short.
static PairShape pair_shape(int b, int n, int m) {
  const int ntiles = (m + kTN - 1) / kTN;
  auto ctas = [&](int warps, int split) { return (long long)b * ((n + warps * kBlk - 1) / (warps * kBlk)) * split; };
  PairShape p;
  p.warps = 8;
  while (p.warps > 2 && ctas(p.warps, ntiles) < 4LL * kNumSMs) p.warps >>= 1;
  p.tiles_per_cta = ntiles;
  while (p.tiles_per_cta > 1 && ctas(p.warps, (ntiles + p.tiles_per_cta - 1) / p.tiles_per_cta) < 12LL * kNumSMs)
    p.tiles_per_cta = (p.tiles_per_cta + 1) / 2;
  p.split = (ntiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.tiles_x = (n + p.warps * kBlk - 1) / (p.warps * kBlk);
  return p;
}

template <int WARPS>
static int pair_launch(int b, int n, int m, const PairShape &sh, int use_tma, const float *xyz1, const float *xyz2,
                       u64 *rowkey, u64 *colkey, cudaStream_t s) {
  const size_t smem = sizeof(PairSmem<WARPS>);
  {
    static size_t granted[kMaxDevices];
    const int rc0 = grant_dyn_smem(chamfer_pair_kernel<WARPS>, smem, granted, 0);
    if (rc0) return rc0;
  }
  chamfer_pair_kernel<WARPS><<<dim3(sh.tiles_x, sh.split, b), WARPS * 32, smem, s>>>(
      n, m, sh.tiles_x, sh.split, sh.tiles_per_cta, use_tma, xyz1, xyz2, rowkey, colkey);
  count_launch();
  return launch_status();
}

int chamfer_fused_launch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1, float *dist2,
                         int *idx1, int *idx2, void *ws, size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < chamfer_fused_workspace_bytes(b, n, m)) return MVP_ERR_WORKSPACE;
  u64 *rowkey = reinterpret_cast<u64 *>(ws);
  u64 *colkey = rowkey + (size_t)b * n;
  {
    cudaError_t e = cudaMemsetAsync(ws, 0xff, chamfer_fused_workspace_bytes(b, n, m), s);
    if (e != cudaSuccess) return (int)e;
  }
  const PairShape sh = pair_shape(b, n, m);
  // bulk TMA needs 16-byte aligned sources: tiles start at multiples of 1024 points (12 KB), clouds at
  // multiples of m*12 bytes
  const int use_tma = ((reinterpret_cast<uintptr_t>(xyz2) & 15) == 0 && (m % 4) == 0) ? 1 : 0;
  int rc;
  if (sh.warps == 8) rc = pair_launch<8>(b, n, m, sh, use_tma, xyz1, xyz2, rowkey, colkey, s);
  else if (sh.warps == 4) rc = pair_launch<4>(b, n, m, sh, use_tma, xyz1, xyz2, rowkey, colkey, s);
  else rc = pair_launch<2>(b, n, m, sh, use_tma, xyz1, xyz2, rowkey, colkey, s);
  if (rc) return rc;

  const int vec2 = ((reinterpret_cast<uintptr_t>(xyz2) & 15) == 0 && (m % 4) == 0) ? 1 : 0;
  const int vec1 = ((reinterpret_cast<uintptr_t>(xyz1) & 15) == 0 && (n % 4) == 0) ? 1 : 0;
  const long long chunks = ((long long)n + m + 7) / 8;
  dim3 rgrid((unsigned)std::min<long long>(chunks, 1 << 20), b);
  chamfer_resolve_kernel<<<rgrid, 256, 0, s>>>(b, n, m, vec1, vec2, xyz1, xyz2, rowkey, colkey, dist1, dist2, idx1, idx2);
  count_launch();
  return launch_status();
}

}  // namespace mvp
