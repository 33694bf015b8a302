from typing import Tuple

import torch
from torch.autograd import Function

from ..._native import _lib


class ThreeInterpolate(Function):
    """Weighted sum of three gathered feature columns — drop-in for the reference's
    utils/mm3d_pn2/ops/interpolate/three_interpolate.py:8-60."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, indices: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        """
        Args:
            features (Tensor): (B, C, M) source features.
            indices (Tensor): (B, n, 3) int32 — three source indices per target.
            weight (Tensor): (B, n, 3) interpolation weights.
        Returns:
            Tensor: (B, C, n).
        """
        assert features.is_contiguous()
        assert indices.is_contiguous()
        assert weight.is_contiguous()
        device = _lib.require_cuda(features, indices, weight, what="three_interpolate")
        if features.dtype != torch.float32 or weight.dtype != torch.float32 or indices.dtype != torch.int32:
            raise TypeError("three_interpolate: features/weight must be float32 and indices int32")
        B, c, m = features.size()
        n = indices.size(1)
        ctx.three_interpolate_for_backward = (indices, weight, m)
        output = torch.empty(B, c, n, device=device, dtype=torch.float32)
        with torch.cuda.device(device):
            rc = _lib.lib.mvp_three_interpolate(B, c, m, n, _lib.ptr(features), _lib.ptr(indices), _lib.ptr(weight),
                                                _lib.ptr(output), _lib.stream_of(features))
        _lib.check(rc, "mvp_three_interpolate")
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """grad_out (B, C, n) -> gradient of features (B, C, M); indices and weights get none."""
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad_out_data = grad_out.data.contiguous()
        grad_features = torch.empty(B, c, m, device=grad_out_data.device, dtype=torch.float32)
        with torch.cuda.device(grad_out_data.device):
            ws = _lib.workspace(_lib.lib.mvp_scatter_workspace_bytes(B, m, 3 * n), grad_out_data.device)
            rc = _lib.lib.mvp_three_interpolate_grad_ws(B, c, n, m, _lib.ptr(grad_out_data), _lib.ptr(idx),
                                                        _lib.ptr(weight), _lib.ptr(grad_features), _lib.ptr(ws),
                                                        ws.numel(), _lib.stream_of(grad_out_data))
        _lib.check(rc, "mvp_three_interpolate_grad")
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply
