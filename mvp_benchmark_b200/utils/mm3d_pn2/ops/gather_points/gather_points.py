import torch
from torch.autograd import Function

from ..._native import _lib


class GatherPoints(Function):
    """out[b, c, m] = features[b, c, indices[b, m]] — drop-in for the reference's
    utils/mm3d_pn2/ops/gather_points/gather_points.py:7-49 (backward: scatter-add)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
        """
        Args:
            features (Tensor): (B, C, N).
            indices (Tensor): (B, M) int32.
        Returns:
            Tensor: (B, C, M).
        """
        assert features.is_contiguous()
        assert indices.is_contiguous()
        device = _lib.require_cuda(features, indices, what="gather_points")
        if features.dtype != torch.float32 or indices.dtype != torch.int32:
            raise TypeError("gather_points: features must be float32 and indices int32")
        B, npoint = indices.size()
        _, C, N = features.size()
        output = torch.empty(B, C, npoint, device=device, dtype=torch.float32)
        with torch.cuda.device(device):
            rc = _lib.lib.mvp_gather_points(B, C, N, npoint, _lib.ptr(features), _lib.ptr(indices),
                                            _lib.ptr(output), _lib.stream_of(features))
        _lib.check(rc, "mvp_gather_points")
        ctx.for_backwards = (indices, C, N)
        ctx.mark_non_differentiable(indices)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_out_data = grad_out.data.contiguous()
        grad_features = torch.empty(B, C, N, device=grad_out_data.device, dtype=torch.float32)
        with torch.cuda.device(grad_out_data.device):
            # scratch for the transposed index (the scatter becomes a sum per element: no float atomics)
            ws = _lib.workspace(_lib.lib.mvp_scatter_workspace_bytes(B, N, npoint), grad_out_data.device)
            rc = _lib.lib.mvp_gather_points_grad_ws(B, C, N, npoint, _lib.ptr(grad_out_data), _lib.ptr(idx),
                                                    _lib.ptr(grad_features), _lib.ptr(ws), ws.numel(),
                                                    _lib.stream_of(grad_out_data))
        _lib.check(rc, "mvp_gather_points_grad")
        return grad_features, None


gather_points = GatherPoints.apply
