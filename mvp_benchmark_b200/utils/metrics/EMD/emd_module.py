"""EMD approximation (parallel auction) — drop-in for the reference's utils/metrics/EMD/emd_module.py
(emdFunction :40-81, emdModule :83-88).

  xyz1, xyz2: (B, n, 3), same n, n % 1024 == 0, B <= 512, coordinates normalised to [0, 1];
  xyz1 is the prediction, xyz2 the ground truth; only xyz1 receives a gradient.
  returns dist (B, n) — squared distance to the assigned target — and assignment (B, n) int32.
  The result is an approximation and the assignment need not be a bijection.

The twelve state tensors the reference allocates per call (:54-65) are one workspace here and the whole
auction is one kernel launch (mvp_emd_forward).  Input errors the reference only printf()s
(emd_cuda.cu:236-249) raise.
"""
import torch
from torch import nn
from torch.autograd import Function

from .._native import _lib


class emdFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        batchsize, n, _ = xyz1.size()
        _, m, _ = xyz2.size()

        assert n == m
        assert xyz1.size()[0] == xyz2.size()[0]
        assert batchsize <= 512

        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.contiguous().float()
        if not xyz1.is_cuda:
            xyz1 = xyz1.cuda()
        if not xyz2.is_cuda:
            xyz2 = xyz2.cuda()
        device = _lib.require_cuda(xyz1, xyz2, dtype=torch.float32, what="emd")
        dist = torch.empty(batchsize, n, device=device, dtype=torch.float32)
        assignment = torch.empty(batchsize, n, device=device, dtype=torch.int32)
        with torch.cuda.device(device):
            ws = _lib.workspace(_lib.lib.mvp_emd_forward_workspace_bytes(batchsize, n), device)
            rc = _lib.lib.mvp_emd_forward(batchsize, n, m, _lib.ptr(xyz1), _lib.ptr(xyz2), float(eps), int(iters),
                                          _lib.ptr(dist), _lib.ptr(assignment), _lib.ptr(ws), ws.numel(),
                                          _lib.stream_of(xyz1))
        _lib.check(rc, "mvp_emd_forward")
        ctx.save_for_backward(xyz1, xyz2, assignment)
        ctx.mark_non_differentiable(assignment)
        return dist, assignment

    @staticmethod
    def backward(ctx, graddist, gradidx=None):
        xyz1, xyz2, assignment = ctx.saved_tensors
        graddist = graddist.contiguous()
        batchsize, n, _ = xyz1.size()
        gradxyz1 = torch.empty_like(xyz1)
        gradxyz2 = torch.zeros_like(xyz2)
        with torch.cuda.device(xyz1.device):
            rc = _lib.lib.mvp_emd_backward(batchsize, n, _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(graddist),
                                           _lib.ptr(assignment), _lib.ptr(gradxyz1), _lib.stream_of(xyz1))
        _lib.check(rc, "mvp_emd_backward")
        return gradxyz1, gradxyz2, None, None


class emdModule(nn.Module):
    def __init__(self):
        super(emdModule, self).__init__()

    def forward(self, input1, input2, eps, iters):
        return emdFunction.apply(input1, input2, eps, iters)
