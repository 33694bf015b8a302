import torch
from torch.autograd import Function

from ..._native import _lib


class KNN(Function):
    """k nearest neighbours by a per-query max-heap (k <= 100) — drop-in for the reference's
    utils/mm3d_pn2/ops/knn/knn.py:7-69."""

    @staticmethod
    def forward(ctx, k: int, xyz: torch.Tensor, center_xyz: torch.Tensor = None,
                transposed: bool = False) -> torch.Tensor:
        """
        Args:
            k (int): number of neighbours.
            xyz (Tensor): (B, N, 3), or (B, 3, N) when `transposed`.
            center_xyz (Tensor): (B, npoint, 3) / (B, 3, npoint) query centres; defaults to xyz.
            transposed (bool): inputs are channel-first.  Pass positionally (knn = KNN.apply).
        Returns:
            Tensor: (B, k, npoint) int32 indices, ascending distance along k.
        """
        assert k > 0
        if center_xyz is None:
            center_xyz = xyz
        if transposed:
            xyz = xyz.transpose(2, 1).contiguous()
            center_xyz = center_xyz.transpose(2, 1).contiguous()
        assert xyz.is_contiguous()
        assert center_xyz.is_contiguous()
        device = _lib.require_cuda(xyz, center_xyz, dtype=torch.float32, what="knn")
        B, npoint, _ = center_xyz.shape
        N = xyz.shape[1]
        idx = torch.empty(B, npoint, k, device=device, dtype=torch.int32)
        dist2 = torch.empty(B, npoint, k, device=device, dtype=torch.float32)
        with torch.cuda.device(device):
            rc = _lib.lib.mvp_knn(B, N, npoint, int(k), _lib.ptr(xyz), _lib.ptr(center_xyz), _lib.ptr(idx),
                                  _lib.ptr(dist2), _lib.stream_of(xyz))
        _lib.check(rc, "mvp_knn")
        idx = idx.transpose(2, 1).contiguous()  # (B, k, npoint), knn.py:63
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


knn = KNN.apply
