import torch


def fscore(dist1, dist2, threshold=0.0001):
    """F-score of two clouds from their (squared) nearest-neighbour distances.

    Same contract as the reference's utils/metrics/CD/fscore.py:3-16: per-cloud fraction of distances
    under `threshold` in each direction, their harmonic mean, NaN (0/0) mapped to 0.
    Returns (fscore, precision_1, precision_2), each of shape (B,).
    """
    p1 = (dist1 < threshold).float().mean(dim=1)
    p2 = (dist2 < threshold).float().mean(dim=1)
    f = 2 * p1 * p2 / (p1 + p2)
    f[torch.isnan(f)] = 0
    return f, p1, p2
