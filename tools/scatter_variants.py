#!/usr/bin/env python
"""Times the backward scatters of gather_points / three_interpolate at VRCNet's shapes: the workspace-free kernels
(shared-memory accumulation) against the workspace (transposed-index) variants.  Run under gpurun."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvp_benchmark_b200 import _lib  # noqa: E402

L, P = _lib.lib, _lib.ptr
dev = torch.device("cuda:0")


def t_ms(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for bb, c, nn, mp in [(64, 64, 3072, 15360), (64, 128, 1536, 7680), (64, 256, 768, 3840), (64, 64, 3072, 2048), (64, 64, 3072, 1536)]:
    go = torch.randn(bb, c, mp, device=dev)
    idx = torch.randint(0, nn, (bb, mp), device=dev, dtype=torch.int32)
    gp, gp2 = torch.empty(bb, c, nn, device=dev), torch.empty(bb, c, nn, device=dev)
    ws = _lib.workspace(L.mvp_scatter_workspace_bytes(bb, nn, mp), dev)
    S = _lib.stream_of(go)
    a = t_ms(lambda: L.mvp_gather_points_grad(bb, c, nn, mp, P(go), P(idx), P(gp), S))
    b = t_ms(lambda: L.mvp_gather_points_grad_ws(bb, c, nn, mp, P(go), P(idx), P(gp2), P(ws), ws.numel(), S))
    print(f"gather_grad {bb}x{c}x{nn}<-{mp}: plain {a:.4f} ms  ws {b:.4f} ms  maxdiff {float((gp - gp2).abs().max()):.2e}", flush=True)
for bb, c, m_, n_ in [(64, 128, 1536, 3072), (64, 256, 768, 1536), (64, 512, 384, 768)]:
    go = torch.randn(bb, c, n_, device=dev)
    i3 = torch.randint(0, m_, (bb, n_, 3), device=dev, dtype=torch.int32)
    w = torch.rand(bb, n_, 3, device=dev)
    gp, gp2 = torch.empty(bb, c, m_, device=dev), torch.empty(bb, c, m_, device=dev)
    ws = _lib.workspace(L.mvp_scatter_workspace_bytes(bb, m_, 3 * n_), dev)
    S = _lib.stream_of(go)
    a = t_ms(lambda: L.mvp_three_interpolate_grad(bb, c, n_, m_, P(go), P(i3), P(w), P(gp), S))
    b = t_ms(lambda: L.mvp_three_interpolate_grad_ws(bb, c, n_, m_, P(go), P(i3), P(w), P(gp2), P(ws), ws.numel(), S))
    print(f"interp_grad {bb}x{c}x{m_}<-{n_}: plain {a:.4f} ms  ws {b:.4f} ms  maxdiff {float((gp - gp2).abs().max()):.2e}", flush=True)
