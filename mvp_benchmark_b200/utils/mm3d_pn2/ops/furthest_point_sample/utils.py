import torch


def calc_square_dist(point_feat_a, point_feat_b, norm=True):
    """Pairwise squared distance |a|^2 + |b|^2 - 2 a.b between (B,N,C) and (B,M,C) -> (B,N,M); with
    `norm` the result is sqrt(.)/C.  Same contract as the reference's
    utils/mm3d_pn2/ops/furthest_point_sample/utils.py:4-31."""
    num_channel = point_feat_a.shape[-1]
    a_square = point_feat_a.pow(2).sum(dim=-1, keepdim=True)       # (B, N, 1)
    b_square = point_feat_b.pow(2).sum(dim=-1).unsqueeze(1)        # (B, 1, M)
    coor = torch.matmul(point_feat_a, point_feat_b.transpose(1, 2))
    dist = a_square + b_square - 2 * coor
    if norm:
        dist = torch.sqrt(dist) / num_channel
    return dist
