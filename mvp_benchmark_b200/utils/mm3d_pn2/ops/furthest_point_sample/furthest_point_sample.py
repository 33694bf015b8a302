import torch
from torch.autograd import Function

from ..._native import _lib


class FurthestPointSampling(Function):
    """Iterative furthest point sampling on coordinates — drop-in for the reference's
    utils/mm3d_pn2/ops/furthest_point_sample/furthest_point_sample.py:7-39.  Starts at index 0; indices
    are bit-identical to the reference kernel, ties included."""

    @staticmethod
    def forward(ctx, points_xyz: torch.Tensor, num_points: int) -> torch.Tensor:
        """
        Args:
            points_xyz (Tensor): (B, N, 3).
            num_points (int): number of samples.
        Returns:
            Tensor: (B, num_points) int32 indices.
        """
        assert points_xyz.is_contiguous()
        device = _lib.require_cuda(points_xyz, dtype=torch.float32, what="furthest_point_sample")
        B, N = points_xyz.size()[:2]
        output = torch.empty(B, num_points, device=device, dtype=torch.int32)
        # The running-distance scratch of the reference (:30) lives in registers; only clouds beyond
        # 32768 points need it in memory.
        temp = torch.empty(B, N, device=device, dtype=torch.float32) if N > 32768 else None
        with torch.cuda.device(device):
            rc = _lib.lib.mvp_furthest_point_sampling(B, N, int(num_points), _lib.ptr(points_xyz), _lib.ptr(temp),
                                                      _lib.ptr(output), _lib.stream_of(points_xyz))
        _lib.check(rc, "mvp_furthest_point_sampling")
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(xyz, a=None):
        return None, None


class FurthestPointSamplingWithDist(Function):
    """Furthest point sampling on a precomputed (B, N, N) distance matrix — drop-in for
    furthest_point_sample.py:42-74."""

    @staticmethod
    def forward(ctx, points_dist: torch.Tensor, num_points: int) -> torch.Tensor:
        """
        Args:
            points_dist (Tensor): (B, N, N) pairwise distances.
            num_points (int): number of samples.
        Returns:
            Tensor: (B, num_points) int32 indices.
        """
        assert points_dist.is_contiguous()
        device = _lib.require_cuda(points_dist, dtype=torch.float32, what="furthest_point_sample_with_dist")
        B, N, _ = points_dist.size()
        output = torch.empty(B, num_points, device=device, dtype=torch.int32)
        temp = torch.empty(B, N, device=device, dtype=torch.float32) if N > 32768 else None
        with torch.cuda.device(device):
            rc = _lib.lib.mvp_furthest_point_sampling_with_dist(B, N, int(num_points), _lib.ptr(points_dist),
                                                                _lib.ptr(temp), _lib.ptr(output),
                                                                _lib.stream_of(points_dist))
        _lib.check(rc, "mvp_furthest_point_sampling_with_dist")
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(xyz, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply
furthest_point_sample_with_dist = FurthestPointSamplingWithDist.apply
