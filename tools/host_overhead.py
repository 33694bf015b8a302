#!/usr/bin/env python
"""Host-side cost of one operator call (tiny inputs, so the GPU is never the bottleneck): the Python autograd
layer vs the raw C-ABI call.  Run under gpurun."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mvp_benchmark_b200  # noqa: E402

mvp_benchmark_b200.install()
import metrics  # noqa: E402
import mm3d_pn2 as mm  # noqa: E402
from mvp_benchmark_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
L, P = _lib.lib, _lib.ptr


def host_us(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    dt = time.perf_counter() - t
    torch.cuda.synchronize()
    return 1e6 * dt / n


x = torch.rand(2, 64, 3, device=dev)
y = torch.rand(2, 64, 3, device=dev)
f = torch.rand(2, 8, 64, device=dev)
idx = torch.randint(0, 64, (2, 32), device=dev, dtype=torch.int32)
out = torch.empty(2, 8, 32, device=dev)
cd = metrics.cd()
s = _lib.stream_of(x)
d1, d2 = torch.empty(2, 64, device=dev), torch.empty(2, 64, device=dev)
i1, i2 = torch.empty(2, 64, device=dev, dtype=torch.int32), torch.empty(2, 64, device=dev, dtype=torch.int32)
ws = _lib.workspace(16, dev)
xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
fg = f.clone().requires_grad_(True)


def cd_fwd_bwd():
    a, b, _, _ = cd(xg, yg)
    (a.sum() + b.sum()).backward()


def gather_fwd_bwd():
    mm.gather_points(fg, idx).sum().backward()


rows = [
    ("raw C ABI mvp_gather_points", lambda: L.mvp_gather_points(2, 8, 64, 32, P(f), P(idx), P(out), s)),
    ("raw C ABI mvp_chamfer_forward", lambda: L.mvp_chamfer_forward(2, 64, 64, P(x), P(y), P(d1), P(d2), P(i1), P(i2), P(ws), 16, s)),
    ("mm3d_pn2.gather_points (no grad)", lambda: mm.gather_points(f, idx)),
    ("mm3d_pn2.furthest_point_sample", lambda: mm.furthest_point_sample(x, 16)),
    ("metrics.cd() forward (no grad)", lambda: cd(x, y)),
    ("mm3d_pn2.gather_points fwd+bwd (+sum)", gather_fwd_bwd),
    ("metrics.cd() fwd+bwd (+2 sums, add)", cd_fwd_bwd),
    ("torch baseline: f.sum().backward()", lambda: fg.sum().backward()),
    ("torch baseline: torch.empty(2,64)", lambda: torch.empty(2, 64, device=dev)),
]
for name, fn in rows:
    print(f"{name:45s} {host_us(fn):8.1f} us/call", flush=True)
