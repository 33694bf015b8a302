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


class _ChamferLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dist1, dist2):
        dist1, dist2 = dist1.contiguous(), dist2.contiguous()
        dev = _lib.require_cuda(dist1, dist2, dtype=torch.float32, what="chamfer_loss")
        if dist1.dim() != 2 or dist2.dim() != 2 or dist1.size(0) != dist2.size(0):
            raise ValueError(f"chamfer_loss: expected (B, N) and (B, M), got {tuple(dist1.shape)} and {tuple(dist2.shape)}")
        B, n = dist1.shape
        m = dist2.size(1)
        cd_p = torch.empty(B, device=dev, dtype=torch.float32)
        cd_t = torch.empty(B, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            rc = _lib.lib.mvp_chamfer_loss(B, n, m, _lib.ptr(dist1), _lib.ptr(dist2), _lib.ptr(cd_p), _lib.ptr(cd_t),
                                           _lib.stream_of(dist1))
        _lib.check(rc, "mvp_chamfer_loss")
        ctx.save_for_backward(dist1, dist2)
        return cd_p, cd_t

    @staticmethod
    def backward(ctx, g_p, g_t):
        dist1, dist2 = ctx.saved_tensors
        B, n = dist1.shape
        m = dist2.size(1)
        g_p = (torch.zeros(B, device=dist1.device) if g_p is None else g_p).contiguous().float()
        g_t = (torch.zeros(B, device=dist1.device) if g_t is None else g_t).contiguous().float()
        gd1, gd2 = torch.empty_like(dist1), torch.empty_like(dist2)
        with torch.cuda.device(dist1.device):
            rc = _lib.lib.mvp_chamfer_loss_grad(B, n, m, _lib.ptr(dist1), _lib.ptr(dist2), _lib.ptr(g_p), _lib.ptr(g_t),
                                                _lib.ptr(gd1), _lib.ptr(gd2), _lib.stream_of(dist1))
        _lib.check(rc, "mvp_chamfer_loss_grad")
        return gd1, gd2


def chamfer_loss(dist1, dist2):
    """(cd_p, cd_t) per cloud from the Chamfer operator's dist1 (B, N) and dist2 (B, M) — the torch glue of
    completion/model_utils.py:71-72 as one reduction kernel (and one elementwise kernel backward):
    cd_p = (sqrt(dist1).mean(1) + sqrt(dist2).mean(1)) / 2,  cd_t = dist1.mean(1) + dist2.mean(1).  Differentiable."""
    return _ChamferLoss.apply(dist1, dist2)


def three_nn_weights(target, source):
    """The three nearest points of `source` (B, M, 3) for every point of `target` (B, N, 3) and their normalised
    inverse-distance weights — completion/model_utils.py:286-293 (three_nn_upsampling: three_nn, clamp, reciprocal,
    sum, divide) as the grid search plus ONE elementwise kernel.  Returns (idx (B, N, 3) int32, weight (B, N, 3));
    not differentiable (the original's weights are not either: three_nn marks its outputs non-differentiable)."""
    target, source = target.detach().contiguous(), source.detach().contiguous()
    dev = _lib.require_cuda(target, source, dtype=torch.float32, what="three_nn_weights")
    B, N, _ = target.shape
    m = source.size(1)
    dist2 = torch.empty(B, N, 3, device=dev, dtype=torch.float32)
    idx = torch.empty(B, N, 3, device=dev, dtype=torch.int32)
    weight = torch.empty(B, N, 3, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        ws = _lib.workspace(_lib.lib.mvp_three_nn_workspace_bytes(B, N, m), dev)
        _lib.check(_lib.lib.mvp_three_nn_ws(B, N, m, _lib.ptr(target), _lib.ptr(source), _lib.ptr(dist2), _lib.ptr(idx),
                                            _lib.ptr(ws), ws.numel(), _lib.stream_of(target)), "mvp_three_nn")
        _lib.check(_lib.lib.mvp_three_nn_weights(B, N, _lib.ptr(dist2), _lib.ptr(weight), _lib.stream_of(target)),
                   "mvp_three_nn_weights")
    return idx, weight
