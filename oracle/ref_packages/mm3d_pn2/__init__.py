"""TEST INFRASTRUCTURE — a `mm3d_pn2` package backed by the REFERENCE's own CUDA kernels
(oracle/_ref/libref_ops.so via oracle/ref_cuda.py) with the reference's autograd structure
(utils/mm3d_pn2/ops/*/*.py).  Only the names the completion models import
(completion/model_utils.py:21, completion/models/vrcnet.py:18, completion/models/ecg.py:19).
Never imported by the product package."""
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle import ref_cuda as _ref  # noqa: E402


class _FPS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points_xyz, num_points):
        assert points_xyz.is_contiguous()
        idx = _ref.furthest_point_sample(points_xyz, num_points)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, g):
        return None, None


class _BallQuery(torch.autograd.Function):
    @staticmethod
    def forward(ctx, min_radius, max_radius, sample_num, xyz, center_xyz):
        assert xyz.is_contiguous() and center_xyz.is_contiguous()
        idx = _ref.ball_query(min_radius, max_radius, sample_num, xyz, center_xyz)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, g):
        return None, None, None, None, None


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, indices):
        assert features.is_contiguous() and indices.is_contiguous()
        ctx.for_backwards = (indices, features.size(2))
        return _ref.gather_points(features, indices)

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.for_backwards
        return _ref.gather_points_grad(grad_out.contiguous(), idx, n), None


class _Group(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, indices):
        assert features.is_contiguous() and indices.is_contiguous()
        ctx.for_backwards = (indices, features.size(2))
        return _ref.group_points(features, indices)

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.for_backwards
        return _ref.group_points_grad(grad_out.contiguous(), idx, n), None


class _ThreeNN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, target, source):
        assert target.is_contiguous() and source.is_contiguous()
        dist2, idx = _ref.three_nn(target, source)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


class _ThreeInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, indices, weight):
        assert features.is_contiguous() and indices.is_contiguous() and weight.is_contiguous()
        ctx.three_interpolate_for_backward = (indices, weight, features.size(2))
        return _ref.three_interpolate(features, indices, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        return _ref.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


furthest_point_sample = _FPS.apply
ball_query = _BallQuery.apply
gather_points = _Gather.apply
grouping_operation = _Group.apply
three_nn = _ThreeNN.apply
three_interpolate = _ThreeInterpolate.apply
