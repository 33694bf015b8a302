from typing import Tuple

import torch
from torch import nn as nn
from torch.autograd import Function

from ..._native import _lib
from ..ball_query import ball_query
from ..knn import knn


class QueryAndGroup(nn.Module):
    """Neighbourhood query (ball query, or kNN when `max_radius` is None) followed by grouping — same
    options and outputs as the reference's utils/mm3d_pn2/ops/group_points/group_points.py:11-122.

    Args:
        max_radius (float | None): ball radius; None selects kNN.
        sample_num (int): neighbours per centre.
        min_radius (float): inner radius. Default 0.
        use_xyz (bool): prepend centred xyz to the grouped features. Default True.
        return_grouped_xyz (bool): also return the grouped xyz. Default False.
        normalize_xyz (bool): divide grouped xyz by max_radius. Default False.
        uniform_sample (bool): resample duplicates uniformly. Default False.
        return_unique_cnt (bool): also return the number of unique neighbours (needs uniform_sample).
    """

    def __init__(self, max_radius, sample_num, min_radius=0, use_xyz=True, return_grouped_xyz=False,
                 normalize_xyz=False, uniform_sample=False, return_unique_cnt=False):
        super(QueryAndGroup, self).__init__()
        self.max_radius = max_radius
        self.min_radius = min_radius
        self.sample_num = sample_num
        self.use_xyz = use_xyz
        self.return_grouped_xyz = return_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.uniform_sample = uniform_sample
        self.return_unique_cnt = return_unique_cnt
        if self.return_unique_cnt:
            assert self.uniform_sample, \
                'uniform_sample should be True when returning the count of unique samples'
        if self.max_radius is None:
            assert not self.normalize_xyz, 'can not normalize grouped xyz when max_radius is None'

    def forward(self, points_xyz, center_xyz, features=None):
        """points_xyz (B, N, 3), center_xyz (B, npoint, 3), features (B, C, N) | None
        -> (B, 3 + C, npoint, sample_num) (plus the optional extras)."""
        if self.max_radius is None:
            idx = knn(self.sample_num, points_xyz, center_xyz, False)
            idx = idx.transpose(1, 2).contiguous()
        else:
            idx = ball_query(self.min_radius, self.max_radius, self.sample_num, points_xyz, center_xyz)

        if self.uniform_sample:
            unique_cnt = torch.zeros((idx.shape[0], idx.shape[1]))
            for i_batch in range(idx.shape[0]):
                for i_region in range(idx.shape[1]):
                    unique_ind = torch.unique(idx[i_batch, i_region, :])
                    num_unique = unique_ind.shape[0]
                    unique_cnt[i_batch, i_region] = num_unique
                    sample_ind = torch.randint(0, num_unique, (self.sample_num - num_unique, ), dtype=torch.long)
                    all_ind = torch.cat((unique_ind, unique_ind[sample_ind]))
                    idx[i_batch, i_region, :] = all_ind

        xyz_trans = points_xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)  # (B, 3, npoint, sample_num)
        grouped_xyz -= center_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz /= self.max_radius

        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, 'Cannot have not features and not use xyz as a feature!'
            new_features = grouped_xyz

        ret = [new_features]
        if self.return_grouped_xyz:
            ret.append(grouped_xyz)
        if self.return_unique_cnt:
            ret.append(unique_cnt)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(nn.Module):
    """Groups every point into one neighbourhood (group_points.py:125-163).

    Args:
        use_xyz (bool): prepend xyz to the features.
    """

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        """xyz (B, N, 3), new_xyz ignored, features (B, C, N) | None -> (B, C + 3, 1, N)."""
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        return grouped_xyz


class GroupingOperation(Function):
    """out[b, c, p, s] = features[b, c, indices[b, p, s]] — drop-in for group_points.py:166-218
    (backward: scatter-add)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
        """
        Args:
            features (Tensor): (B, C, N).
            indices (Tensor): (B, npoint, nsample) int32.
        Returns:
            Tensor: (B, C, npoint, nsample).
        """
        assert features.is_contiguous()
        assert indices.is_contiguous()
        device = _lib.require_cuda(features, indices, what="grouping_operation")
        if features.dtype != torch.float32 or indices.dtype != torch.int32:
            raise TypeError("grouping_operation: features must be float32 and indices int32")
        B, nfeatures, nsample = indices.size()
        _, C, N = features.size()
        output = torch.empty(B, C, nfeatures, nsample, device=device, dtype=torch.float32)
        with torch.cuda.device(device):
            rc = _lib.lib.mvp_group_points(B, C, N, nfeatures, nsample, _lib.ptr(features), _lib.ptr(indices),
                                           _lib.ptr(output), _lib.stream_of(features))
        _lib.check(rc, "mvp_group_points")
        ctx.for_backwards = (indices, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """grad_out (B, C, npoint, nsample) -> gradient of features (B, C, N)."""
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad_out_data = grad_out.data.contiguous()
        grad_features = torch.empty(B, C, N, device=grad_out_data.device, dtype=torch.float32)
        with torch.cuda.device(grad_out_data.device):
            ws = _lib.workspace(_lib.lib.mvp_scatter_workspace_bytes(B, N, npoint * nsample), grad_out_data.device)
            rc = _lib.lib.mvp_group_points_grad_ws(B, C, N, npoint, nsample, _lib.ptr(grad_out_data), _lib.ptr(idx),
                                                   _lib.ptr(grad_features), _lib.ptr(ws), ws.numel(),
                                                   _lib.stream_of(grad_out_data))
        _lib.check(rc, "mvp_group_points_grad")
        return grad_features, None


grouping_operation = GroupingOperation.apply
