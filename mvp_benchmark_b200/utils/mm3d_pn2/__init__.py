"""Drop-in for the reference's utils/mm3d_pn2/__init__.py:1-20.

Exports the point-cloud ops of the hot path.  The reference also re-exports 2-D detection ops from the
un-vendored `mmcv` (nms, RoIAlign, roi_align, sigmoid_focal_loss, SigmoidFocalLoss) and
NaiveSyncBatchNorm1d/2d; none of them is used by completion/  and they are out
of scope here: they resolve lazily to mmcv's own objects when mmcv is installed and raise ImportError
otherwise, so `import mm3d_pn2` itself never needs mmcv.
"""
from .ops import (ball_query, knn, furthest_point_sample, furthest_point_sample_with_dist,
                  three_interpolate, three_nn, gather_points, grouping_operation, group_points, GroupAll,
                  QueryAndGroup, get_compiler_version, get_compiling_cuda_version, Points_Sampler)
from . import ops as _ops

__all__ = [
    'ball_query', 'knn', 'furthest_point_sample', 'furthest_point_sample_with_dist', 'three_interpolate',
    'three_nn', 'gather_points', 'grouping_operation', 'group_points', 'GroupAll', 'QueryAndGroup',
    'get_compiler_version', 'get_compiling_cuda_version', 'Points_Sampler',
]


def __getattr__(name):
    return getattr(_ops, name)
