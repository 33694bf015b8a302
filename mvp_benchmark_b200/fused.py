"""Fused operators beyond the reference's operator packages: what the completion models do
with torch glue around the hot path, as single exact CUDA searches.  Opt-in — see model_patches.py."""
import torch

from . import _lib


def knn_points(k, cloud, queries=None):
    """k nearest points of `cloud` (B, M, 3) for every point of `queries` (B, N, 3; default: the cloud itself,
    so every point's first neighbour is itself).  Returns (dist2 (B, N, k) float32 squared distances,
    idx (B, N, k) int32), ascending in (distance, index).  Not differentiable (indices; recompute distances
    from the gathered points where a gradient is needed — model_patches.knn_point does).
    Replaces the (B, N, M) matmul + torch.topk of completion/model_utils.py:242-259."""
    if queries is None:
        queries = cloud
    cloud = cloud.detach().contiguous()
    queries = queries.detach().contiguous()
    dev = _lib.require_cuda(queries, cloud, dtype=torch.float32, what="knn_points")
    if cloud.dim() != 3 or queries.dim() != 3 or cloud.size(2) != 3 or queries.size(2) != 3:
        raise ValueError(f"knn_points: expected (B, N, 3) tensors, got {tuple(queries.shape)} and {tuple(cloud.shape)}")
    B, N, _ = queries.shape
    M = cloud.size(1)
    if cloud.size(0) != B:
        raise ValueError("knn_points: batch sizes differ")
    k = int(k)
    if not 1 <= k <= min(M, 64):
        raise ValueError(f"knn_points: k={k} out of range 1..min(M, 64) (M={M})")
    dist2 = torch.empty(B, N, k, device=dev, dtype=torch.float32)
    idx = torch.empty(B, N, k, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        ws = _lib.workspace(_lib.lib.mvp_knn_points_workspace_bytes(B, N, M, k), dev)
        rc = _lib.lib.mvp_knn_points(B, N, M, k, _lib.ptr(queries), _lib.ptr(cloud), _lib.ptr(dist2), _lib.ptr(idx),
                                     _lib.ptr(ws), ws.numel(), _lib.stream_of(queries))
    _lib.check(rc, "mvp_knn_points")
    return dist2, idx
