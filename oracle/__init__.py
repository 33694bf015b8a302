"""TEST INFRASTRUCTURE — numpy-facing wrapper of the CPU oracle (oracle/oracle.c) and, when it was
built, of the reference CUDA kernels recompiled for sm_100a (oracle/_ref/libref_ops.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package.  Nothing under mvp_benchmark_b200/ does: the product path has no CPU fallback.

Each function names the reference lines it restates in oracle/oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """gcc recipe for liboracle.so (plain C + OpenMP; -mavx2 -mfma so fmaf is one instruction)."""
    src = os.path.join(_HERE, "oracle.c")
    if not force and os.path.isfile(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src):
        return _LIB_PATH
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-mavx2", "-mfma",
           "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", src, "-o", _LIB_PATH, "-lm"]
    subprocess.check_call(cmd)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_emd_forward.restype = ctypes.c_int
        _lib.oracle_emd_forward_state.restype = ctypes.c_int
        _lib.oracle_knn.restype = ctypes.c_int
        _lib.oracle_fps_block_size.restype = ctypes.c_int
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def chamfer_forward(xyz1, xyz2):
    """chamfer3D.cu:12-154 -> (dist1, dist2, idx1, idx2)."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist1 = np.empty((b, n), np.float32)
    dist2 = np.empty((b, m), np.float32)
    idx1 = np.empty((b, n), np.int32)
    idx2 = np.empty((b, m), np.int32)
    lib().oracle_chamfer_forward(b, n, m, _fp(xyz1), _fp(xyz2), _fp(dist1), _fp(dist2), _ip(idx1), _ip(idx2))
    return dist1, dist2, idx1, idx2


def chamfer_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2):
    """chamfer3D.cu:155-195 -> (gradxyz1, gradxyz2)."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    g1, g2, idx1, idx2 = _f32(graddist1), _f32(graddist2), _i32(idx1), _i32(idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1 = np.empty((b, n, 3), np.float32)
    gx2 = np.empty((b, m, 3), np.float32)
    lib().oracle_chamfer_backward(b, n, m, _fp(xyz1), _fp(xyz2), _fp(g1), _fp(g2), _ip(idx1), _ip(idx2),
                                  _fp(gx1), _fp(gx2))
    return gx1, gx2


def chamfer_sample(xyz1, xyz2, graddist1, graddist2, nq):
    """Bounded CPU sample of one Chamfer forward+backward for ONE cloud pair (n,3)/(m,3): both directions
    for the first `nq` queries of each cloud against the whole other cloud (chamfer3D.cu:12-134), then the
    backward scatter of those queries (:155-174).  Pair evaluations = nq*m + nq*n.  Returns
    (dist1[:nq], idx1[:nq], dist2[:nq], idx2[:nq], gradxyz1, gradxyz2)."""
    xyz1, xyz2, g1, g2 = _f32(xyz1), _f32(xyz2), _f32(graddist1), _f32(graddist2)
    n, m = xyz1.shape[0], xyz2.shape[0]
    q1, q2 = min(nq, n), min(nq, m)
    d1, i1 = np.empty(q1, np.float32), np.empty(q1, np.int32)
    d2, i2 = np.empty(q2, np.float32), np.empty(q2, np.int32)
    L = lib()
    L.oracle_nm_distance(q1, _fp(xyz1), m, _fp(xyz2), _fp(d1), _ip(i1))
    L.oracle_nm_distance(q2, _fp(xyz2), n, _fp(xyz1), _fp(d2), _ip(i2))
    gx1, gx2 = np.zeros((n, 3), np.float32), np.zeros((m, 3), np.float32)
    L.oracle_nm_distance_grad(q1, _fp(xyz1), m, _fp(xyz2), _fp(g1), _ip(i1), _fp(gx1), _fp(gx2))
    L.oracle_nm_distance_grad(q2, _fp(xyz2), n, _fp(xyz1), _fp(g2), _ip(i2), _fp(gx2), _fp(gx1))
    return d1, i1, d2, i2, gx1, gx2


def emd_forward(xyz1, xyz2, eps, iters, return_price=False):
    """emd_cuda.cu:23-282 -> (dist, assignment[, price])."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    assert xyz2.shape == xyz1.shape
    dist = np.empty((b, n), np.float32)
    asg = np.empty((b, n), np.int32)
    price = np.empty((b, n), np.float32)
    rc = lib().oracle_emd_forward(b, n, _fp(xyz1), _fp(xyz2), ctypes.c_float(eps), int(iters), _fp(dist),
                                  _ip(asg), _fp(price))
    if rc != 0:
        raise ValueError("oracle_emd_forward: invalid input (b<=512, n%1024==0 required)")
    return (dist, asg, price) if return_price else (dist, asg)


def emd_last_ambiguous():
    """Number of GetMax windows (target, round) that held two or more bidders during the last
    emd_forward / emd_forward_state call: on such inputs the reference's result is a last-writer race
    (emd_cuda.cu:188-191) and no deterministic implementation can be required to reproduce it."""
    lib().oracle_emd_last_ambiguous.restype = ctypes.c_longlong
    return int(lib().oracle_emd_last_ambiguous())


def emd_last_ambiguous_per_cloud(b):
    """The same count for each of the `b` clouds of the last emd_forward call (int64 array)."""
    out = np.zeros(b, np.int64)
    lib().oracle_emd_last_ambiguous_per_cloud(out.ctypes.data_as(ctypes.c_void_p), int(b))
    return out


def emd_set_tie_policy(p):
    """0 = highest source index wins GetMax near-ties (default, = the product), 1 = lowest."""
    lib().oracle_emd_set_tie_policy(int(p))


def emd_forward_state(xyz1, xyz2, eps, iters):
    """One cloud (n,3): dict of every state array after `iters` rounds (debug aid for the GetMax race)."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    n = xyz1.shape[0]
    f = lambda: np.empty(n, np.float32)  # noqa: E731
    i = lambda: np.empty(n, np.int32)  # noqa: E731
    last_unass = i()
    st = dict(dist=f(), assignment=i(), price=f(), assignment_inv=i(), bid=i(), bid_increments=f(),
              max_increments=f(), max_idx=i())
    rc = lib().oracle_emd_forward_state(n, _fp(xyz1), _fp(xyz2), ctypes.c_float(eps), int(iters), _fp(st["dist"]),
                                        _ip(st["assignment"]), _fp(st["price"]), _ip(st["assignment_inv"]),
                                        _ip(st["bid"]), _fp(st["bid_increments"]), _fp(st["max_increments"]),
                                        _ip(st["max_idx"]), _ip(last_unass))
    if rc < 0:
        raise ValueError("oracle_emd_forward_state: n % 1024 == 0 required")
    st["last_unassigned"] = last_unass[:rc].copy()
    return st


def emd_backward(xyz1, xyz2, graddist, assignment):
    """emd_cuda.cu:284-316 -> gradxyz1."""
    xyz1, xyz2, g, a = _f32(xyz1), _f32(xyz2), _f32(graddist), _i32(assignment)
    b, n, _ = xyz1.shape
    out = np.empty((b, n, 3), np.float32)
    lib().oracle_emd_backward(b, n, _fp(xyz1), _fp(xyz2), _fp(g), _ip(a), _fp(out))
    return out


def fps_block_size(n):
    """furthest_point_sample_cuda.cu:11-15."""
    return lib().oracle_fps_block_size(int(n))


def furthest_point_sample(xyz, m):
    """furthest_point_sample_cuda.cu:26-141 -> idx (B, m) int32."""
    xyz = _f32(xyz)
    b, n, _ = xyz.shape
    idx = np.zeros((b, m), np.int32)
    lib().oracle_fps(b, n, int(m), _fp(xyz), _ip(idx))
    return idx


def furthest_point_sample_with_dist(dist, m):
    """furthest_point_sample_cuda.cu:214-331 -> idx (B, m) int32."""
    dist = _f32(dist)
    b, n, _ = dist.shape
    idx = np.zeros((b, m), np.int32)
    lib().oracle_fps_with_dist(b, n, int(m), _fp(dist), _ip(idx))
    return idx


def ball_query(min_radius, max_radius, nsample, xyz, center_xyz):
    """ball_query_cuda.cu:11-54 -> idx (B, npoint, nsample) int32."""
    xyz, c = _f32(xyz), _f32(center_xyz)
    b, n, _ = xyz.shape
    m = c.shape[1]
    idx = np.empty((b, m, nsample), np.int32)
    lib().oracle_ball_query(b, n, m, ctypes.c_float(min_radius), ctypes.c_float(max_radius), int(nsample),
                            _fp(c), _fp(xyz), _ip(idx))
    return idx


def gather_points(points, idx):
    """gather_points_cuda.cu:8-26 -> (B, C, M)."""
    points, idx = _f32(points), _i32(idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = np.empty((b, c, m), np.float32)
    lib().oracle_gather_points(b, c, n, m, _fp(points), _ip(idx), _fp(out))
    return out


def gather_points_grad(grad_out, idx, n):
    """gather_points_cuda.cu:51-70 -> (B, C, N)."""
    grad_out, idx = _f32(grad_out), _i32(idx)
    b, c, m = grad_out.shape
    out = np.empty((b, c, n), np.float32)
    lib().oracle_gather_points_grad(b, c, n, m, _fp(grad_out), _ip(idx), _fp(out))
    return out


def group_points(points, idx):
    """group_points_cuda.cu:56-79 -> (B, C, npoint, nsample)."""
    points, idx = _f32(points), _i32(idx)
    b, c, n = points.shape
    _, p, s = idx.shape
    out = np.empty((b, c, p, s), np.float32)
    lib().oracle_group_points(b, c, n, p, s, _fp(points), _ip(idx), _fp(out))
    return out


def group_points_grad(grad_out, idx, n):
    """group_points_cuda.cu:10-31 -> (B, C, N)."""
    grad_out, idx = _f32(grad_out), _i32(idx)
    b, c, p, s = grad_out.shape
    out = np.empty((b, c, n), np.float32)
    lib().oracle_group_points_grad(b, c, n, p, s, _fp(grad_out), _ip(idx), _fp(out))
    return out


def three_nn(unknown, known):
    """three_nn_cuda.cu:11-65 -> (dist2 SQUARED (B,N,3), idx (B,N,3))."""
    unknown, known = _f32(unknown), _f32(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    d = np.empty((b, n, 3), np.float32)
    idx = np.empty((b, n, 3), np.int32)
    lib().oracle_three_nn(b, n, m, _fp(unknown), _fp(known), _fp(d), _ip(idx))
    return d, idx


def three_interpolate(points, idx, weight):
    """three_interpolate_cuda.cu:11-35 -> (B, C, n)."""
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out = np.empty((b, c, n), np.float32)
    lib().oracle_three_interpolate(b, c, m, n, _fp(points), _ip(idx), _fp(weight), _fp(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """three_interpolate_cuda.cu:61-84 -> (B, C, m)."""
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    b, c, n = grad_out.shape
    out = np.empty((b, c, m), np.float32)
    lib().oracle_three_interpolate_grad(b, c, n, m, _fp(grad_out), _ip(idx), _fp(weight), _fp(out))
    return out


def knn(k, xyz, center_xyz):
    """knn_cuda.cu:58-94 -> (idx (B, npoint, k) int32, dist2 (B, npoint, k)) — before the Python
    transpose of knn.py:63."""
    xyz, c = _f32(xyz), _f32(center_xyz)
    b, n, _ = xyz.shape
    m = c.shape[1]
    idx = np.empty((b, m, k), np.int32)
    d = np.empty((b, m, k), np.float32)
    rc = lib().oracle_knn(b, n, m, int(k), _fp(xyz), _fp(c), _ip(idx), _fp(d))
    if rc != 0:
        raise ValueError("oracle_knn: 0 < k <= 100 required (knn_cuda.cu:72-73)")
    return idx, d


def knn_points(k, cloud, queries=None):
    """completion/model_utils.py:242-259 (`knn`, `knn_point`) as an exact search: (dist2 (B, N, k), idx (B, N, k)
    int32) ascending in (distance, index); queries default to the cloud itself."""
    c = _f32(cloud)
    q = c if queries is None else _f32(queries)
    b, n, _ = q.shape
    m = c.shape[1]
    idx = np.empty((b, n, k), np.int32)
    d = np.empty((b, n, k), np.float32)
    rc = lib().oracle_knn_points(b, n, m, int(k), _fp(q), _fp(c), _fp(d), _ip(idx))
    if rc != 0:
        raise ValueError("oracle_knn_points: 1 <= k <= m required")
    return d, idx


def chamfer_loss(dist1, dist2):
    """calc_cd's epilogue (completion/model_utils.py:71-72) -> (cd_p (B,), cd_t (B,))."""
    d1, d2 = _f32(dist1), _f32(dist2)
    b, n = d1.shape
    m = d2.shape[1]
    cd_p, cd_t = np.empty(b, np.float32), np.empty(b, np.float32)
    if lib().oracle_chamfer_loss(b, n, m, _fp(d1), _fp(d2), _fp(cd_p), _fp(cd_t)) != 0:
        raise ValueError("oracle_chamfer_loss: n, m > 0 required")
    return cd_p, cd_t


def pointwise_conv(x, weight, bias=None, relu=False):
    """A 1x1 convolution as the completion models' nn.Conv1d / nn.Conv2d(kernel_size=1) layers compute it
    (completion/models/vrcnet.py:26-38,68,161-164,197-207; ecg.py:46; pcn.py encoder): x (B, C, N), weight (O, C),
    bias (O) or None -> (B, O, N); products and sums in float64, rounded once to fp32 (the exact contraction: a fp32 or
    TF32 kernel is compared with a tolerance, small-integer operands exactly)."""
    y = np.einsum("oc,bcn->bon", np.asarray(weight, np.float64), np.asarray(x, np.float64))
    if bias is not None:
        y = y + np.asarray(bias, np.float64)[None, :, None]
    if relu:
        y = np.maximum(y, 0.0)
    return y.astype(np.float32)


def pointwise_conv_grads(x, weight, grad_out):
    """The same layer's gradients: (grad_x (B, C, N), grad_w (O, C), grad_bias (O)), float64 sums rounded once."""
    g, w, x64 = np.asarray(grad_out, np.float64), np.asarray(weight, np.float64), np.asarray(x, np.float64)
    return (np.einsum("oc,bon->bcn", w, g).astype(np.float32), np.einsum("bon,bcn->oc", g, x64).astype(np.float32),
            g.sum((0, 2)).astype(np.float32))


def max_last(x):
    """`torch.max(x, -1)` as completion/models/ecg.py:64 and completion/model_utils.py:53,104 use it: (values, the FIRST
    position of the maximum), a NaN winning over any number (the first NaN)."""
    a = np.asarray(x, np.float32)
    nan = np.isnan(a)
    arg = np.where(nan.any(-1), nan.argmax(-1), np.where(nan, -np.inf, a).argmax(-1))
    return np.take_along_axis(a, arg[..., None], -1)[..., 0], arg.astype(np.int64)


def fscore(dist1, dist2, threshold=0.0001, mean="factor"):
    """utils/metrics/CD/fscore.py:12-15 -> (fscore, precision_1, precision_2), each (B,), in torch's fp32 arithmetic:
    fscore = ((2 p1) p2) / (p1 + p2), NaN -> 0.  The mean of the 0/1 values is count * fl(1 / n) the way torch's CUDA
    reduce kernel takes it (mean="factor": the sum times a float factor) or count / n the way its CPU path does
    (mean="div": sum, then div_) — one ulp apart when n is not a power of two."""
    d1, d2 = _f32(dist1), _f32(dist2)
    one = np.float32(1.0)
    c1 = (d1 < np.float32(threshold)).sum(1).astype(np.float32)
    c2 = (d2 < np.float32(threshold)).sum(1).astype(np.float32)
    if mean == "factor":
        p1, p2 = c1 * (one / np.float32(d1.shape[1])), c2 * (one / np.float32(d2.shape[1]))
    else:
        p1, p2 = c1 / np.float32(d1.shape[1]), c2 / np.float32(d2.shape[1])
    with np.errstate(invalid="ignore", divide="ignore"):
        f = (np.float32(2.0) * p1 * p2) / (p1 + p2)
    f[np.isnan(f)] = 0
    return f.astype(np.float32), p1.astype(np.float32), p2.astype(np.float32)
