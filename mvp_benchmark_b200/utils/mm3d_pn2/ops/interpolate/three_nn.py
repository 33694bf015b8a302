from typing import Tuple

import torch
from torch.autograd import Function

from ..._native import _lib


class ThreeNN(Function):
    """Three nearest source points of every target point — drop-in for the reference's
    utils/mm3d_pn2/ops/interpolate/three_nn.py:8-42."""

    @staticmethod
    def forward(ctx, target: torch.Tensor, source: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """
        Args:
            target (Tensor): (B, N, 3) points that look for neighbours.
            source (Tensor): (B, M, 3) points searched.
        Returns:
            (Tensor, Tensor): (B, N, 3) L2 distances (square-rooted, ascending) and (B, N, 3) int32 indices.
        """
        assert target.is_contiguous()
        assert source.is_contiguous()
        device = _lib.require_cuda(target, source, dtype=torch.float32, what="three_nn")
        B, N, _ = target.size()
        m = source.size(1)
        dist2 = torch.empty(B, N, 3, device=device, dtype=torch.float32)
        idx = torch.empty(B, N, 3, device=device, dtype=torch.int32)
        with torch.cuda.device(device):
            # scratch for the grid over `source` (exact search, same outputs as the exhaustive kernel)
            ws = _lib.workspace(_lib.lib.mvp_three_nn_workspace_bytes(B, N, m), device)
            rc = _lib.lib.mvp_three_nn_ws(B, N, m, _lib.ptr(target), _lib.ptr(source), _lib.ptr(dist2),
                                          _lib.ptr(idx), _lib.ptr(ws), ws.numel(), _lib.stream_of(target))
        _lib.check(rc, "mvp_three_nn")
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply
