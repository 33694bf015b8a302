import torch
from torch.autograd import Function

from ..._native import _lib


class BallQuery(Function):
    """Ball query — drop-in for the reference's utils/mm3d_pn2/ops/ball_query/ball_query.py:7-44.

    For every centre, the first `sample_num` points (ascending index) whose squared distance d2 satisfies
    d2 == 0 or min_radius^2 <= d2 < max_radius^2; short rows are padded with the first hit, rows without
    any hit are all zero.
    """

    @staticmethod
    def forward(ctx, min_radius: float, max_radius: float, sample_num: int, xyz: torch.Tensor,
                center_xyz: torch.Tensor) -> torch.Tensor:
        """
        Args:
            min_radius, max_radius (float): radii of the shell.
            sample_num (int): maximum number of points per ball.
            xyz (Tensor): (B, N, 3) points.
            center_xyz (Tensor): (B, npoint, 3) ball centres.
        Returns:
            Tensor: (B, npoint, sample_num) int32 indices.
        """
        assert center_xyz.is_contiguous()
        assert xyz.is_contiguous()
        assert min_radius < max_radius
        device = _lib.require_cuda(xyz, center_xyz, dtype=torch.float32, what="ball_query")
        B, N, _ = xyz.size()
        npoint = center_xyz.size(1)
        idx = torch.empty(B, npoint, sample_num, device=device, dtype=torch.int32)
        with torch.cuda.device(device):
            rc = _lib.lib.mvp_ball_query(B, N, npoint, float(min_radius), float(max_radius), int(sample_num),
                                         _lib.ptr(center_xyz), _lib.ptr(xyz), _lib.ptr(idx), _lib.stream_of(xyz))
        _lib.check(rc, "mvp_ball_query")
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None, None


ball_query = BallQuery.apply
