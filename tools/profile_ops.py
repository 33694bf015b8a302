#!/usr/bin/env python
"""One call of every operator family at the sizes the completion models use — the subject of the per-kernel ncu
captures summarised under profiles/ (run it under `ncu --set full -k regex:mvp`; keep it short: ncu replays every
kernel ~40 times).   python tools/profile_ops.py [family ...]   families: fps gather group interp three_nn knn_points
ball_query knn fused pointwise chamfer_brute chamfer_bwd emd"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mvp_benchmark_b200  # noqa: E402
mvp_benchmark_b200.install()
import metrics  # noqa: E402
import mm3d_pn2 as mm  # noqa: E402
from mvp_benchmark_b200 import _lib, fused  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
R = lambda *s: torch.rand(*s, device=dev, generator=g)  # noqa: E731
N = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
want = set(sys.argv[1:])
on = lambda name: not want or name in want  # noqa: E731

if on("fps"):
    for b, n, m in ((32, 2048, 2048), (64, 3072, 1536), (16, 8192, 2048)):
        mm.furthest_point_sample(R(b, n, 3), m)
    fused.fps_gather(R(64, 1536, 3), 768)
if on("gather"):
    f = N(64, 64, 3072).requires_grad_(True)
    i = torch.randint(0, 3072, (64, 15360), device=dev, generator=g, dtype=torch.int32)
    mm.gather_points(f, i).backward(N(64, 64, 15360))
if on("group"):
    f = N(64, 128, 1536).requires_grad_(True)
    i = torch.randint(0, 1536, (64, 768, 1), device=dev, generator=g, dtype=torch.int32)
    mm.grouping_operation(f, i).backward(N(64, 128, 768, 1))
if on("three_nn") or on("interp"):
    tgt, src = R(64, 3072, 3), R(64, 1536, 3)
    idx, w = fused.three_nn_weights(tgt, src)
    f = N(64, 128, 1536).requires_grad_(True)
    mm.three_interpolate(f, idx, w).backward(N(64, 128, 3072))
if on("knn_points"):
    x = R(64, 3072, 3)
    fused.knn_points(16, x)
    fused.knn_points(10, x, R(64, 1536, 3))
if on("ball_query"):
    pcd = R(32, 2048, 3)
    fused.ball_query_group(0, 0.0632455532, 8, pcd, pcd[:, :102].contiguous())
    mm.ball_query(0, 0.1095445115, 24, pcd, pcd[:, :102].contiguous())
if on("knn"):
    mm.knn(16, R(64, 2048, 3), R(64, 512, 3), False)
    mm.knn(64, R(16, 8192, 3), R(16, 2048, 3), False)
if on("pointwise"):
    # the tcgen05 1x1 layer: resident-weight kernel (64 -> 256 and 128 -> 256: forward, input gradient), streaming kernel
    # (512 -> 512), and the bias kernels of the wide layers
    for (b, c, o, n) in ((64, 64, 256, 3072), (64, 256, 64, 3072), (64, 128, 256, 2048), (64, 512, 512, 384)):
        fused._pointwise_conv_raw(N(b, c, n), N(o, c), N(o))
    fused._pointwise_wgrad_raw(N(64, 256, 3072), N(64, 64, 3072), True)
    fused._pointwise_wgrad_raw(N(64, 16, 3072), N(64, 64, 3072), True)
    y = N(64, 1024, 2048)
    fused.bias_add_(y, N(1024))
    fused.channel_sum(y)
if on("fused"):
    f = N(64, 64, 3072).requires_grad_(True)
    i = torch.randint(0, 3072, (64, 1536, 10), device=dev, generator=g, dtype=torch.int32)
    fused.gather_max(f, i).sum().backward()
    y = N(64, 16, 3072).requires_grad_(True)
    w = N(64, 2, 20, 3072).requires_grad_(True)
    fused.neighbor_weighted_sum(y, torch.randint(0, 3072, (64, 3072, 20), device=dev, generator=g, dtype=torch.int32), w).sum().backward()
    fused.topk_rows(N(32, 2048, 2048), 16)
if on("chamfer_brute"):
    b, n = 32, 16384
    a, c = R(b, n, 3), R(b, n, 3)
    outs = [torch.empty(b, n, device=dev), torch.empty(b, n, device=dev), torch.empty(b, n, device=dev, dtype=torch.int32),
            torch.empty(b, n, device=dev, dtype=torch.int32)]
    ws = _lib.workspace(_lib.lib.mvp_chamfer_forward_workspace_bytes(b, n, n), dev)
    _lib.check(_lib.lib.mvp_chamfer_forward_algo(1, b, n, n, _lib.ptr(a), _lib.ptr(c), *[_lib.ptr(o) for o in outs], _lib.ptr(ws),
                                                 ws.numel(), _lib.stream_of(a)), "brute")
if on("chamfer_bwd"):
    a, c = R(32, 16384, 3).requires_grad_(True), R(32, 16384, 3).requires_grad_(True)
    d1, d2, _, _ = metrics.cd()(a, c)
    torch.autograd.backward([d1, d2], [R(32, 16384), R(32, 16384)])
if on("emd"):
    metrics.emd()(R(64, 8192, 3), R(64, 8192, 3), 0.005, 50)
torch.cuda.synchronize()
print("done")
