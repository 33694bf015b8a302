#!/usr/bin/env python
"""Generates tests/golden/ref_cuda_golden.npz from the REFERENCE's own CUDA kernels
(oracle/_ref/libref_ops.so = the reference .cu files compiled unmodified for sm_100a) on a B200:

    gpurun -- 'python tests/golden/make_golden_gpu.py'      # writes gpurun_out/ref_cuda_golden.npz
    cp gpurun_out/ref_cuda_golden.npz tests/golden/

These vectors pin the CPU oracle (tests/test_oracle_golden.py, no GPU needed) for every op the
reference's own tests leave unpinned.  Inputs are stored next to the outputs.
EMD is run twice; `*_stable` records whether the reference agreed with itself (its GetMax step is a
last-writer race, emd_cuda.cu:188-191).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _data  # noqa: E402
from oracle import ref_cuda  # noqa: E402

dev = torch.device("cuda:0")
out = {}


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def put(name, **kw):
    for k, v in kw.items():
        out[f"{name}.{k}"] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)


rng = np.random.default_rng(1234)

# ---- Chamfer forward + backward
for name, kind, b, n, m, seed in [("cd_uniform", "uniform", 2, 300, 500, 0), ("cd_lattice", "lattice", 2, 600, 520, 1),
                                  ("cd_dups", "duplicates", 2, 1100, 257, 2), ("cd_tiny", "uniform", 3, 1, 5, 3)]:
    x1, x2 = _data.cloud(kind, b, n, seed), _data.cloud(kind, b, m, seed + 100)
    g1, g2 = rng.random((b, n), dtype=np.float32), rng.random((b, m), dtype=np.float32)
    d1, d2, i1, i2 = ref_cuda.chamfer_forward(T(x1), T(x2))
    gx1, gx2 = ref_cuda.chamfer_backward(T(x1), T(x2), T(g1), T(g2), i1, i2)
    put(name, xyz1=x1, xyz2=x2, dist1=d1, dist2=d2, idx1=i1, idx2=i2, graddist1=g1, graddist2=g2, gradxyz1=gx1,
        gradxyz2=gx2)

# ---- EMD
for name, kind, b, n, eps, iters, seed in [("emd_train", "uniform", 2, 1024, 0.005, 50, 0),
                                           ("emd_short", "sphere", 2, 1024, 0.005, 3, 1),
                                           ("emd_2048", "uniform", 1, 2048, 0.004, 120, 2),
                                           ("emd_3072", "uniform", 1, 3072, 0.005, 20, 3)]:
    x1, x2 = _data.cloud(kind, b, n, seed), _data.cloud(kind, b, n, seed + 100)
    d, a = ref_cuda.emd_forward(T(x1), T(x2), eps, iters)
    dd, aa = ref_cuda.emd_forward(T(x1), T(x2), eps, iters)
    g = rng.random((b, n), dtype=np.float32)
    gx = ref_cuda.emd_backward(T(x1), T(x2), T(g), a)
    put(name, xyz1=x1, xyz2=x2, eps=np.float32(eps), iters=np.int32(iters), dist=d, assignment=a,
        stable=np.bool_(bool((a == aa).all().item())), graddist=g, gradxyz1=gx)

# ---- FPS (block size T = 1024 / 512 / 256 / 64 paths, ties through lattice + duplicates, m == n)
for name, kind, b, n, m, seed in [("fps_uniform", "uniform", 2, 1000, 200, 0), ("fps_2048", "uniform", 1, 2048, 512, 1),
                                  ("fps_lattice", "lattice", 2, 700, 700, 2), ("fps_dups", "duplicates", 2, 384, 384, 3),
                                  ("fps_small", "uniform", 2, 100, 100, 4), ("fps_3072", "sphere", 1, 3072, 300, 5)]:
    x = _data.cloud(kind, b, n, seed)
    put(name, xyz=x, m=np.int32(m), idx=ref_cuda.furthest_point_sample(T(x), m))
x = _data.cloud("uniform", 2, 200, 7)
dm = ((x[:, :, None, :] - x[:, None, :, :]) ** 2).sum(-1).astype(np.float32)
put("fpsd", dist=dm, m=np.int32(64), idx=ref_cuda.furthest_point_sample_with_dist(T(dm), 64))

# ---- ball_query (ECG-like radii, completion/model_utils.py:201-211) + inner radius + lattice
for name, kind, b, n, p, rmin, rmax, ns, seed in [("bq_ecg", "uniform", 2, 1024, 51, 0.0, 0.0632455532, 4, 0),
                                                  ("bq_wide", "uniform", 2, 1500, 77, 0.0, 0.2, 24, 1),
                                                  ("bq_shell", "uniform", 2, 800, 40, 0.05, 0.15, 16, 2),
                                                  ("bq_lattice", "lattice", 2, 600, 64, 0.0, 0.125, 12, 3),
                                                  ("bq_none", "uniform", 1, 300, 20, 0.0, 1e-4, 5, 4)]:
    x = _data.cloud(kind, b, n, seed)
    c = _data.cloud(kind, b, p, seed + 50) if name != "bq_ecg" else x[:, :p].copy()
    put(name, xyz=x, centers=c, min_radius=np.float32(rmin), max_radius=np.float32(rmax), nsample=np.int32(ns),
        idx=ref_cuda.ball_query(rmin, rmax, ns, T(x), T(c)))

# ---- gather / group with backward
feat = rng.standard_normal((2, 7, 300)).astype(np.float32)
gi = rng.integers(0, 300, (2, 130)).astype(np.int32)
go = rng.standard_normal((2, 7, 130)).astype(np.float32)
put("gather", points=feat, idx=gi, out=ref_cuda.gather_points(T(feat), T(gi)), grad_out=go,
    grad_points=ref_cuda.gather_points_grad(T(go), T(gi), 300))
gi3 = rng.integers(0, 300, (2, 40, 6)).astype(np.int32)
go3 = rng.standard_normal((2, 7, 40, 6)).astype(np.float32)
put("group", points=feat, idx=gi3, out=ref_cuda.group_points(T(feat), T(gi3)), grad_out=go3,
    grad_points=ref_cuda.group_points_grad(T(go3), T(gi3), 300))

# ---- three_nn / three_interpolate
for name, kind, b, n, m, seed in [("nn3_uniform", "uniform", 2, 300, 100, 0), ("nn3_lattice", "lattice", 2, 200, 150, 1),
                                  ("nn3_two", "uniform", 1, 50, 2, 2)]:
    u, k = _data.cloud(kind, b, n, seed), _data.cloud(kind, b, m, seed + 9)
    d, i = ref_cuda.three_nn(T(u), T(k))
    put(name, unknown=u, known=k, dist2=d, idx=i)
pts = rng.standard_normal((2, 9, 100)).astype(np.float32)
ii = rng.integers(0, 100, (2, 300, 3)).astype(np.int32)
w = rng.random((2, 300, 3), dtype=np.float32)
go = rng.standard_normal((2, 9, 300)).astype(np.float32)
put("interp", points=pts, idx=ii, weight=w, out=ref_cuda.three_interpolate(T(pts), T(ii), T(w)), grad_out=go,
    grad_points=ref_cuda.three_interpolate_grad(T(go), T(ii), T(w), 100))

# ---- knn
for name, kind, b, n, p, k, seed in [("knn_uniform", "uniform", 2, 500, 100, 5, 0), ("knn_lattice", "lattice", 2, 300, 64, 16, 1),
                                     ("knn_big_k", "uniform", 1, 150, 30, 100, 2), ("knn_k_gt_n", "uniform", 1, 7, 9, 10, 3)]:
    x, c = _data.cloud(kind, b, n, seed), _data.cloud(kind, b, p, seed + 77)
    i, d = ref_cuda.knn(k, T(x), T(c))
    put(name, xyz=x, centers=c, k=np.int32(k), idx=i, dist2=d)

torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "ref_cuda_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")
print("emd stable flags:", {k: bool(v) for k, v in out.items() if k.endswith(".stable")})
