#!/usr/bin/env python
"""Timing aid: Chamfer forward / backward, three_nn and knn_points at the VRCNet sizes (B = 64) and the headline size,
through the C ABI with preallocated outputs (CUDA events, median of 20)."""
import ctypes
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvp_benchmark_b200 import _lib

L, P = _lib.lib, _lib.ptr
dev = torch.device("cuda:0")


def med(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2] * 1e3


def graph_med(fn, iters=20):
    """device time without host launch overhead: the call captured into a CUDA graph, replayed"""
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(); fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
    torch.cuda.synchronize()
    return med(g.replay, iters)


S = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
for b, n, m in ((64, 2048, 1024), (64, 2048, 2048), (64, 2048, 3072), (32, 16384, 1024), (32, 16384, 16384)):
    x1, x2 = torch.rand(b, n, 3, device=dev), torch.rand(b, m, 3, device=dev)
    d1, d2 = torch.empty(b, n, device=dev), torch.empty(b, m, device=dev)
    i1, i2 = torch.empty(b, n, device=dev, dtype=torch.int32), torch.empty(b, m, device=dev, dtype=torch.int32)
    g1, g2 = torch.rand(b, n, device=dev), torch.rand(b, m, device=dev)
    o1, o2 = torch.empty_like(x1), torch.empty_like(x2)
    ws = _lib.workspace(L.mvp_chamfer_forward_workspace_bytes(b, n, m), dev)
    fwd = lambda: L.mvp_chamfer_forward(b, n, m, P(x1), P(x2), P(d1), P(d2), P(i1), P(i2), P(ws), ws.numel(), S())  # noqa: E731
    bwd = lambda: L.mvp_chamfer_backward(b, n, m, P(x1), P(x2), P(g1), P(g2), P(i1), P(i2), P(o1), P(o2), S())  # noqa: E731
    print(f"chamfer {b}x{n}x{m}: fwd {med(fwd):7.1f} us  bwd {med(bwd):6.1f} us   graph-replayed: fwd {graph_med(fwd):7.1f} bwd {graph_med(bwd):6.1f}", flush=True)
for b, n, m in ((64, 3072, 1536), (64, 1536, 768), (64, 768, 384)):
    u, k = torch.rand(b, n, 3, device=dev), torch.rand(b, m, 3, device=dev)
    d, i = torch.empty(b, n, 3, device=dev), torch.empty(b, n, 3, device=dev, dtype=torch.int32)
    ws = _lib.workspace(L.mvp_three_nn_workspace_bytes(b, n, m), dev)
    print(f"three_nn {b}x{n} from {m}: {med(lambda: L.mvp_three_nn_ws(b, n, m, P(u), P(k), P(d), P(i), P(ws), ws.numel(), S())):7.1f} us")
for b, n, m, k in ((64, 3072, 3072, 16), (64, 1536, 1536, 16), (64, 1536, 3072, 10), (64, 384, 768, 10)):
    c, q = torch.rand(b, m, 3, device=dev), torch.rand(b, n, 3, device=dev)
    d, i = torch.empty(b, n, k, device=dev), torch.empty(b, n, k, device=dev, dtype=torch.int32)
    ws = _lib.workspace(L.mvp_knn_points_workspace_bytes(b, n, m, k), dev)
    print(f"knn_points {b}x{n} from {m} k={k}: {med(lambda: L.mvp_knn_points(b, n, m, k, P(q), P(c), P(d), P(i), P(ws), ws.numel(), S())):7.1f} us")
