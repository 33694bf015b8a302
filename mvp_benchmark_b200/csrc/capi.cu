// Library-level entry points of libmvp_ops.so: identification, error text, launch counter.
#include "common.cuh"

namespace mvp {
unsigned long long g_launch_count = 0;
}

MVP_API int mvp_abi_version(void) { return 1; }

MVP_API const char *mvp_build_info(void) {
  return "libmvp_ops sm_100a (compute_100a) nvcc " __DATE__ " " __TIME__;
}

MVP_API unsigned long long mvp_launch_count(void) { return mvp::g_launch_count; }

MVP_API const char *mvp_error_string(int code) {
  switch (code) {
    case MVP_OK: return "ok";
    case MVP_ERR_INVALID_ARGUMENT: return "invalid argument (negative size, null pointer or unsupported shape)";
    case MVP_ERR_EMD_SIZE_MISMATCH: return "EMD: the two point clouds should have the same size";
    case MVP_ERR_EMD_BATCH: return "EMD: the batch size should be no greater than 512";
    case MVP_ERR_EMD_MULTIPLE_1024: return "EMD: the size of the point clouds should be a multiple of 1024";
    case MVP_ERR_WORKSPACE: return "workspace missing or too small";
    case MVP_ERR_UNSUPPORTED_DEVICE: return "device is not sm_100";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}
