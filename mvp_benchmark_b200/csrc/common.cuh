// Shared device/host helpers for libmvp_ops.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/mvp_ops.h"

#define MVP_API extern "C" __attribute__((visibility("default")))

namespace mvp {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

extern unsigned long long g_launch_count;  // defined in capi.cu
inline void count_launch(int n = 1) {  // callers may be DataParallel worker threads, one per device
  __atomic_fetch_add(&g_launch_count, (unsigned long long)n, __ATOMIC_RELAXED);
}

inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MVP_OK : (int)e;
}

// Opt a kernel in to `bytes` of dynamic shared memory (above the default 48 KB).  The attribute belongs to the CURRENT
// DEVICE, and the reference drives these operators from nn.DataParallel — one process, one thread per device
// (completion/train.py:49) — so the high-water mark is kept per device: `granted` is a zero-initialised static array
// of kMaxDevices entries owned by the call site.  Concurrent callers on one device set the same value: benign.
constexpr int kMaxDevices = 64;
template <typename K>
inline int grant_dyn_smem(K kernel, size_t bytes, size_t *granted, size_t preset = 40 * 1024) {
  if (bytes <= preset) return MVP_OK;  // static + dynamic share the default 48 KB: nothing to opt in to
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  const bool tracked = dev >= 0 && dev < kMaxDevices;
  if (tracked && bytes <= granted[dev]) return MVP_OK;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return (int)e;
  if (tracked) granted[dev] = bytes;
  return MVP_OK;
}

// The reference's squared distance `dx*dx + dy*dy + dz*dz` as nvcc contracts it (SASS-verified for
// chamfer3D.cu:35, furthest_point_sample_cuda.cu:65-66, ball_query_cuda.cu:41-42, three_nn_cuda.cu:41,
// knn_cuda.cu:82, emd_cuda.cu:146,225):  t = dy*dy;  t = fma(dx,dx,t);  d = fma(dz,dz,t).
__device__ __forceinline__ float sqdist(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Inverse-distance weights of the three nearest sources (completion/model_utils.py:286-293: sqrt of three_nn's squared
// distances, clamp at 1e-10, reciprocal, sum over the three, divide), every operation the IEEE one torch performs, in
// the order torch 2.11's reduction adds three elements (first + third, then second: measured, tests pin it).
__device__ __forceinline__ void three_nn_weights_of(float d0, float d1, float d2, float *w) {
  const float r0 = __fdiv_rn(1.0f, fmaxf(__fsqrt_rn(d0), 1e-10f));
  const float r1 = __fdiv_rn(1.0f, fmaxf(__fsqrt_rn(d1), 1e-10f));
  const float r2 = __fdiv_rn(1.0f, fmaxf(__fsqrt_rn(d2), 1e-10f));
  const float norm = __fadd_rn(__fadd_rn(r0, r2), r1);
  w[0] = __fdiv_rn(r0, norm), w[1] = __fdiv_rn(r1, norm), w[2] = __fdiv_rn(r2, norm);
}

__device__ __forceinline__ uint32_t redux_min_u32(uint32_t v) {
  uint32_t r;
  asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ uint32_t redux_max_u32(uint32_t v) {
  uint32_t r;
  asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}

__device__ __forceinline__ float ld_nc(const float *p) { return __ldg(p); }

// Programmatic dependent launch (the chain of small dependent kernels of the Chamfer forward): a kernel launched with
// launch_pdl() may be scheduled while its predecessor in the stream is still running; it must call pdl_wait() before it
// touches anything the predecessor wrote (returns once the predecessor's grid has completed and its writes are
// visible).  pdl_trigger() in the predecessor says "my dependents may be scheduled from here on" — placed at the top,
// so that the dependent's CTAs fill the slots the predecessor's last wave leaves free and its launch latency
// disappears behind the predecessor's tail.  Both are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace mvp
