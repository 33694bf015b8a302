"""TEST INFRASTRUCTURE — a `metrics` package backed by the REFERENCE's own CUDA kernels
(oracle/_ref/libref_ops.so via oracle/ref_cuda.py), with the reference's autograd structure
(utils/metrics/CD/chamfer3D/dist_chamfer_3D.py:26-74, utils/metrics/EMD/emd_module.py:40-88,
utils/metrics/CD/fscore.py:3-16).  Put `oracle/ref_packages` first on sys.path and the reference's
completion models run on the reference kernels — the "reference CUDA ops on the same B200" arm of
tools/model_step.py.  Never imported by the product package."""
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle import ref_cuda as _ref  # noqa: E402


class _ChamferFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = _ref.chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        return _ref.chamfer_backward(xyz1, xyz2, graddist1.contiguous(), graddist2.contiguous(), idx1, idx2)


class cd(torch.nn.Module):
    def forward(self, input1, input2):
        return _ChamferFunction.apply(input1.contiguous(), input2.contiguous())


class _EmdFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        xyz1, xyz2 = xyz1.contiguous().float(), xyz2.contiguous().float()
        dist, assignment = _ref.emd_forward(xyz1, xyz2, eps, iters)
        ctx.save_for_backward(xyz1, xyz2, assignment)
        return dist, assignment

    @staticmethod
    def backward(ctx, graddist, gradidx):
        xyz1, xyz2, assignment = ctx.saved_tensors
        return _ref.emd_backward(xyz1, xyz2, graddist.contiguous(), assignment), torch.zeros_like(xyz2), None, None


class emd(torch.nn.Module):
    def forward(self, input1, input2, eps, iters):
        return _EmdFunction.apply(input1, input2, eps, iters)


def fscore(dist1, dist2, threshold=0.0001):
    p1 = torch.mean((dist1 < threshold).float(), dim=1)
    p2 = torch.mean((dist2 < threshold).float(), dim=1)
    f = 2 * p1 * p2 / (p1 + p2)
    f[torch.isnan(f)] = 0
    return f, p1, p2
