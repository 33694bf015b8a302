"""Points_Sampler and its D-FPS / F-FPS / FS strategies — same behaviour as the reference's
utils/mm3d_pn2/ops/furthest_point_sample/points_sampler.py:34-158, without the mmcv dependency
(`force_fp32` there only casts half inputs to float, :2,:68)."""
from typing import List

import torch
from torch import nn as nn

from .furthest_point_sample import (furthest_point_sample,
                                    furthest_point_sample_with_dist)
from .utils import calc_square_dist


def _fp32(t):
    return t.float() if isinstance(t, torch.Tensor) and t.dtype in (torch.float16, torch.bfloat16) else t


def get_sampler_type(sampler_type):
    """'D-FPS' | 'F-FPS' | 'FS' -> sampler class."""
    table = {'D-FPS': DFPS_Sampler, 'F-FPS': FFPS_Sampler, 'FS': FS_Sampler}
    if sampler_type not in table:
        raise ValueError('Only "sampler_type" of "D-FPS", "F-FPS", or "FS"'
                         f' are supported, got {sampler_type}')
    return table[sampler_type]


class Points_Sampler(nn.Module):
    """Applies one FPS strategy per range of the input points and concatenates the indices.

    Args:
        num_point (list[int]): samples per range.
        fps_mod_list (list[str]): 'D-FPS' (coordinates), 'F-FPS' (features) or 'FS' (both) per range.
        fps_sample_range_list (list[int]): end of each range, -1 = to the end.
    """

    def __init__(self, num_point: List[int], fps_mod_list: List[str] = ['D-FPS'],
                 fps_sample_range_list: List[int] = [-1]):
        super(Points_Sampler, self).__init__()
        assert len(num_point) == len(fps_mod_list) == len(fps_sample_range_list)
        self.num_point = num_point
        self.fps_sample_range_list = fps_sample_range_list
        self.samplers = nn.ModuleList()
        for fps_mod in fps_mod_list:
            self.samplers.append(get_sampler_type(fps_mod)())
        self.fp16_enabled = False

    def forward(self, points_xyz, features):
        """points_xyz (B, N, 3), features (B, C, N) or None -> (B, sum(num_point)) int32 indices."""
        points_xyz, features = _fp32(points_xyz), _fp32(features)
        indices = []
        last_fps_end_index = 0
        for fps_sample_range, sampler, npoint in zip(self.fps_sample_range_list, self.samplers, self.num_point):
            assert fps_sample_range < points_xyz.shape[1]
            if fps_sample_range == -1:
                sample_points_xyz = points_xyz[:, last_fps_end_index:]
                sample_features = features[:, :, last_fps_end_index:] if features is not None else None
            else:
                sample_points_xyz = points_xyz[:, last_fps_end_index:fps_sample_range]
                sample_features = features[:, :, last_fps_end_index:fps_sample_range] \
                    if features is not None else None
            fps_idx = sampler(sample_points_xyz.contiguous(), sample_features, npoint)
            indices.append(fps_idx + last_fps_end_index)
            last_fps_end_index += fps_sample_range
        return torch.cat(indices, dim=1)


class DFPS_Sampler(nn.Module):
    """FPS on Euclidean point distances."""

    def forward(self, points, features, npoint):
        return furthest_point_sample(points.contiguous(), npoint)


class FFPS_Sampler(nn.Module):
    """FPS on distances in (xyz ++ feature) space."""

    def forward(self, points, features, npoint):
        assert features is not None, 'feature input to FFPS_Sampler should not be None'
        features_for_fps = torch.cat([points, features.transpose(1, 2)], dim=2)
        features_dist = calc_square_dist(features_for_fps, features_for_fps, norm=False)
        return furthest_point_sample_with_dist(features_dist.contiguous(), npoint)


class FS_Sampler(nn.Module):
    """F-FPS and D-FPS side by side."""

    def forward(self, points, features, npoint):
        assert features is not None, 'feature input to FS_Sampler should not be None'
        features_for_fps = torch.cat([points, features.transpose(1, 2)], dim=2)
        features_dist = calc_square_dist(features_for_fps, features_for_fps, norm=False)
        fps_idx_ffps = furthest_point_sample_with_dist(features_dist.contiguous(), npoint)
        fps_idx_dfps = furthest_point_sample(points.contiguous(), npoint)
        return torch.cat([fps_idx_ffps, fps_idx_dfps], dim=1)
