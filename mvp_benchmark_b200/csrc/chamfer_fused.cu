// Large-cloud Chamfer fast path (placeholder until the fused kernel lands): reports "unsupported" so
// mvp_chamfer_forward uses the generic one-direction kernel of chamfer.cu.
#include "common.cuh"

namespace mvp {
bool chamfer_fused_supported(int, int, int) { return false; }
size_t chamfer_fused_workspace_bytes(int, int, int) { return 16; }
int chamfer_fused_launch(int, int, int, const float *, const float *, float *, float *, int *, int *,
                         void *, size_t, cudaStream_t) {
  return MVP_ERR_INVALID_ARGUMENT;
}
}  // namespace mvp
