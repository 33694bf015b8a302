"""Drop-in for the reference's utils/mm3d_pn2/ops/__init__.py (hot-path part)."""
from .ball_query import ball_query
from .furthest_point_sample import (Points_Sampler, furthest_point_sample,
                                    furthest_point_sample_with_dist)
from .gather_points import gather_points
from .group_points import (GroupAll, QueryAndGroup, group_points,
                           grouping_operation)
from .interpolate import three_interpolate, three_nn
from .knn import knn

__all__ = [
    'ball_query', 'knn', 'furthest_point_sample', 'furthest_point_sample_with_dist', 'three_interpolate',
    'three_nn', 'gather_points', 'grouping_operation', 'group_points', 'GroupAll', 'QueryAndGroup',
    'get_compiler_version', 'get_compiling_cuda_version', 'Points_Sampler',
]

_MMCV_PASSTHROUGH = ('nms', 'RoIAlign', 'roi_align', 'sigmoid_focal_loss', 'SigmoidFocalLoss')
_OUT_OF_SCOPE = ('NaiveSyncBatchNorm1d', 'NaiveSyncBatchNorm2d')


def get_compiling_cuda_version():
    """CUDA toolkit libmvp_ops.so was compiled with (the reference forwards mmcv's helper, ops/__init__.py:1-3)."""
    from .._native import _lib
    return _lib.lib.mvp_build_info().decode()


def get_compiler_version():
    from .._native import _lib
    return _lib.lib.mvp_build_info().decode()


def __getattr__(name):
    if name in _MMCV_PASSTHROUGH:
        try:
            import mmcv.ops as _mo
        except ImportError as e:
            raise ImportError(f"mm3d_pn2.{name} is a pass-through to mmcv.ops.{name}; mmcv is not installed") from e
        return getattr(_mo, name)
    if name in _OUT_OF_SCOPE:
        raise ImportError(f"mm3d_pn2.{name} is outside the point-cloud operator hot path and is not provided")
    raise AttributeError(name)
