"""TEST INFRASTRUCTURE — torch-tensor-facing wrapper of oracle/_ref/libref_ops.so: the REFERENCE's own
CUDA kernels, compiled unmodified for sm_100a by oracle/build_ref.py.  Used by the `-m gpu` parity
tests as the ground truth and by bench.py as the "reference CUDA ops on the same B200" timing leg.
Never imported by the product package.

Every function allocates and pre-initialises its outputs exactly the way the reference's Python does
(cited per function) and launches on the current stream (Chamfer/EMD: legacy default stream, as the
reference hard-codes).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_ops.so")
_lib = None


def available():
    return os.path.isfile(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing — run `python oracle/build_ref.py` where /root/reference exists")
        _lib = ctypes.CDLL(LIB_PATH)
        for name in ("ref_chamfer_forward", "ref_chamfer_backward", "ref_emd_forward", "ref_emd_backward"):
            getattr(_lib, name).restype = ctypes.c_int
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f(v):
    return ctypes.c_float(v)


def chamfer_forward(xyz1, xyz2):
    """dist_chamfer_3D.py:28-47 -> (dist1, dist2, idx1, idx2)."""
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    dist1 = torch.zeros(b, n, device=dev)
    dist2 = torch.zeros(b, m, device=dev)
    idx1 = torch.zeros(b, n, device=dev, dtype=torch.int32)
    idx2 = torch.zeros(b, m, device=dev, dtype=torch.int32)
    lib().ref_chamfer_forward(b, n, m, _p(xyz1), _p(xyz2), _p(dist1), _p(dist2), _p(idx1), _p(idx2))
    return dist1, dist2, idx1, idx2


def chamfer_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2):
    """dist_chamfer_3D.py:50-64 -> (gradxyz1, gradxyz2)."""
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = torch.zeros_like(xyz1)
    g2 = torch.zeros_like(xyz2)
    gd1, gd2 = graddist1.contiguous(), graddist2.contiguous()
    lib().ref_chamfer_backward(b, n, m, _p(xyz1), _p(xyz2), _p(g1), _p(g2), _p(gd1), _p(gd2), _p(idx1), _p(idx2))
    return g1, g2


def emd_forward(xyz1, xyz2, eps, iters, return_state=False):
    """emd_module.py:42-70 -> (dist, assignment) [or the dict of all state tensors]."""
    b, n, _ = xyz1.shape
    dev = xyz1.device
    i32 = dict(device=dev, dtype=torch.int32)
    dist = torch.zeros(b, n, device=dev)
    assignment = torch.zeros(b, n, **i32) - 1
    assignment_inv = torch.zeros(b, n, **i32) - 1
    price = torch.zeros(b, n, device=dev)
    bid = torch.zeros(b, n, **i32)
    bid_increments = torch.zeros(b, n, device=dev)
    max_increments = torch.zeros(b, n, device=dev)
    unass_idx = torch.zeros(b * n, **i32)
    max_idx = torch.zeros(b * n, **i32)
    unass_cnt = torch.zeros(512, **i32)
    unass_cnt_sum = torch.zeros(512, **i32)
    cnt_tmp = torch.zeros(512, **i32)
    lib().ref_emd_forward(b, n, _p(xyz1), _p(xyz2), _p(dist), _p(assignment), _p(price), _p(assignment_inv),
                          _p(bid), _p(bid_increments), _p(max_increments), _p(unass_idx), _p(unass_cnt),
                          _p(unass_cnt_sum), _p(cnt_tmp), _p(max_idx), _f(eps), int(iters))
    if return_state:
        return dict(dist=dist, assignment=assignment, price=price, assignment_inv=assignment_inv, bid=bid,
                    bid_increments=bid_increments, max_increments=max_increments, max_idx=max_idx.view(b, n))
    return dist, assignment


def emd_backward(xyz1, xyz2, graddist, assignment):
    """emd_module.py:73-81 -> gradxyz1."""
    b, n, _ = xyz1.shape
    g = torch.zeros_like(xyz1)
    gd = graddist.contiguous()
    lib().ref_emd_backward(b, n, _p(xyz1), _p(xyz2), _p(g), _p(gd), _p(assignment))
    return g


def furthest_point_sample(xyz, m):
    """furthest_point_sample.py:15-35."""
    b, n, _ = xyz.shape
    out = torch.zeros(b, m, device=xyz.device, dtype=torch.int32)
    temp = torch.full((b, n), 1e10, device=xyz.device)
    lib().ref_fps(b, n, int(m), _p(xyz), _p(temp), _p(out), _s())
    return out


def furthest_point_sample_with_dist(dist, m):
    """furthest_point_sample.py:50-70."""
    b, n, _ = dist.shape
    out = torch.zeros(b, m, device=dist.device, dtype=torch.int32)
    temp = torch.full((b, n), 1e10, device=dist.device)
    lib().ref_fps_with_dist(b, n, int(m), _p(dist), _p(temp), _p(out), _s())
    return out


def ball_query(min_radius, max_radius, nsample, xyz, center_xyz):
    """ball_query.py:15-40."""
    b, n, _ = xyz.shape
    m = center_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, device=xyz.device, dtype=torch.int32)
    lib().ref_ball_query(b, n, m, _f(min_radius), _f(max_radius), int(nsample), _p(center_xyz), _p(xyz), _p(idx), _s())
    return idx


def gather_points(points, idx):
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.empty(b, c, m, device=points.device)
    lib().ref_gather_points(b, c, n, m, _p(points), _p(idx), _p(out), _s())
    return out


def gather_points_grad(grad_out, idx, n):
    b, c, m = grad_out.shape
    out = torch.zeros(b, c, n, device=grad_out.device)
    lib().ref_gather_points_grad(b, c, n, m, _p(grad_out), _p(idx), _p(out), _s())
    return out


def group_points(points, idx):
    b, c, n = points.shape
    _, p, s = idx.shape
    out = torch.empty(b, c, p, s, device=points.device)
    lib().ref_group_points(b, c, n, p, s, _p(points), _p(idx), _p(out), _s())
    return out


def group_points_grad(grad_out, idx, n):
    b, c, p, s = grad_out.shape
    out = torch.zeros(b, c, n, device=grad_out.device)
    lib().ref_group_points_grad(b, c, n, p, s, _p(grad_out), _p(idx), _p(out), _s())
    return out


def three_nn(unknown, known):
    """three_nn.py:26-38 — returns SQUARED distances (before the Python sqrt) and idx."""
    b, n, _ = unknown.shape
    m = known.shape[1]
    d = torch.empty(b, n, 3, device=unknown.device)
    idx = torch.empty(b, n, 3, device=unknown.device, dtype=torch.int32)
    lib().ref_three_nn(b, n, m, _p(unknown), _p(known), _p(d), _p(idx), _s())
    return d, idx


def three_interpolate(points, idx, weight):
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.empty(b, c, n, device=points.device)
    lib().ref_three_interpolate(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out), _s())
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    b, c, n = grad_out.shape
    out = torch.zeros(b, c, m, device=grad_out.device)
    lib().ref_three_interpolate_grad(b, c, n, m, _p(grad_out), _p(idx), _p(weight), _p(out), _s())
    return out


def knn(k, xyz, center_xyz):
    """knn.py:57-62 — (idx (B, npoint, k), dist2) before the Python transpose."""
    b, n, _ = xyz.shape
    m = center_xyz.shape[1]
    idx = torch.zeros(b, m, k, device=xyz.device, dtype=torch.int32)
    d = torch.zeros(b, m, k, device=xyz.device)
    lib().ref_knn(b, n, m, int(k), _p(xyz), _p(center_xyz), _p(idx), _p(d), _s())
    return idx, d
