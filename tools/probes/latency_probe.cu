// Latency probe for the FPS selection chain (tools/probes; not part of the library): dependent chains of the
// primitives one pick is made of, cycles per link measured with clock64() inside one CTA.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probes/latency_probe tools/probes/latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int kIters = 2048;

__device__ __forceinline__ int redux_max_s32(int v) {
  int r;
  asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ uint32_t redux_min_u32(uint32_t v) {
  uint32_t r;
  asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ float redux_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// mode: 0 redux.max.s32   1 redux.max.s32 -> compare -> redux.min.u32 (the warp-level pair)   2 shfl butterfly max (5 steps)
//       3 redux.max.f32   4 ballot + ffs + shfl (find the lane holding the max, fetch its key)   5 dependent LDS
//       6 bar.sync only   7 STS slot -> bar.sync -> LDS slot[lane] -> 2 redux (today's second level)
//       8 STS slot -> bar.sync -> NW x LDS + compare tree (no second-level redux)   9 fmnmx chain (ALU reference)
template <int mode>
__global__ void probe(long long *out, int *sink) {
  __shared__ int s_val[2][32];
  __shared__ uint32_t s_key[2][32];
  __shared__ int s_chain[1024];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x / 32;
  for (int i = tid; i < 1024; i += blockDim.x) s_chain[i] = (i * 37 + 11) & 1023;
  if (tid < 32) s_val[0][tid] = s_val[1][tid] = (int)0x80000000, s_key[0][tid] = s_key[1][tid] = 0xffffffffu;
  __syncthreads();
  int v = tid * 2654435 + 12345;
  uint32_t key = tid;
  float f = (float)tid;
  const long long t0 = clock64();
  for (int j = 0; j < kIters; j++) {
    if (mode == 0) {
      v = redux_max_s32(v) + lane - j;
    } else if (mode == 1) {
      const int w = redux_max_s32(v);
      const uint32_t k = redux_min_u32(v == w ? key : 0xffffffffu);
      v = (int)k + lane - j;
    } else if (mode == 2) {
#pragma unroll
      for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
      v += lane - j;
    } else if (mode == 3) {
      f = redux_max_f32(f) + (float)lane - (float)j;
    } else if (mode == 4) {
      const int w = redux_max_s32(v);
      const unsigned b = __ballot_sync(0xffffffffu, v == w);
      const uint32_t k = __shfl_sync(0xffffffffu, key, __ffs(b) - 1);
      v = (int)k + lane - j;
    } else if (mode == 5) {
      v = s_chain[v & 1023];
    } else if (mode == 6) {
      __syncthreads();
    } else if (mode == 7) {
      const int buf = j & 1;
      if (lane == 0) s_val[buf][warp] = v, s_key[buf][warp] = key;
      __syncthreads();
      const int sv = s_val[buf][lane];
      const uint32_t sk = s_key[buf][lane];
      const int bv = redux_max_s32(sv);
      const uint32_t bk = redux_min_u32(sv == bv ? sk : 0xffffffffu);
      v = (int)bk + tid - j;
    } else if (mode == 8) {
      const int buf = j & 1;
      if (lane == 0) s_val[buf][warp] = v, s_key[buf][warp] = key;
      __syncthreads();
      int bv = (int)0x80000000;
      uint32_t bk = 0xffffffffu;
      for (int w4 = 0; w4 < nw; w4 += 4) {
        const int4 a = *reinterpret_cast<const int4 *>(&s_val[buf][w4]);
        const uint4 k4 = *reinterpret_cast<const uint4 *>(&s_key[buf][w4]);
        const int av[4] = {a.x, a.y, a.z, a.w};
        const uint32_t ak[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (av[e] > bv || (av[e] == bv && ak[e] < bk)) bv = av[e], bk = ak[e];
      }
      v = (int)bk + tid - j;
    } else {
      f = fmaxf(f * 1.0001f, (float)j);
    }
  }
  const long long t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
  sink[tid] = v + (int)f;
}

int main() {
  long long *out;
  int *sink;
  cudaMalloc(&out, 8);
  cudaMalloc(&sink, 4096);
  const char *names[] = {"redux.max.s32", "redux.max + redux.min", "shfl butterfly x5", "redux.max.f32",
                         "redux.max + ballot + shfl", "LDS dependent", "bar.sync", "STS+bar+LDS+2 redux",
                         "STS+bar+LDS.128 tree", "fmul+fmax (ALU)"};
  auto run = [&](int mode, int threads) {
    for (int rep = 0; rep < 2; rep++) {
      switch (mode) {
        case 0: probe<0><<<1, threads>>>(out, sink); break;
        case 1: probe<1><<<1, threads>>>(out, sink); break;
        case 2: probe<2><<<1, threads>>>(out, sink); break;
        case 3: probe<3><<<1, threads>>>(out, sink); break;
        case 4: probe<4><<<1, threads>>>(out, sink); break;
        case 5: probe<5><<<1, threads>>>(out, sink); break;
        case 6: probe<6><<<1, threads>>>(out, sink); break;
        case 7: probe<7><<<1, threads>>>(out, sink); break;
        case 8: probe<8><<<1, threads>>>(out, sink); break;
        default: probe<9><<<1, threads>>>(out, sink); break;
      }
    }
    long long c;
    cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    printf("%-28s threads %4d : %7.1f cycles/link\n", names[mode], threads, (double)c / kIters);
  };
  for (int mode = 0; mode < 10; mode++)
    for (int threads : {32, 128, 256, 512, 1024}) {
      if (mode != 6 && mode != 7 && mode != 8 && threads != 32 && threads != 256) continue;
      run(mode, threads);
    }
  return cudaDeviceSynchronize() != cudaSuccess;
}
