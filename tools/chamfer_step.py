#!/usr/bin/env python
"""A few Chamfer forward + backward steps through the C ABI at the headline size (for ncu launch lists and
`--set full` captures: keep it short).  python tools/chamfer_step.py [--algo auto|brute|grid|grid_thread] [--steps 3]
[--kind uniform] [--b 32 --n 16384 --m 16384]"""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _data  # noqa: E402
from mvp_benchmark_b200 import _lib as L  # noqa: E402

ALGO = {"auto": 0, "brute": 1, "grid": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--kind", default="uniform")
    ap.add_argument("--kind2", default=None)
    ap.add_argument("--b", type=int, default=32)
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--m", type=int, default=16384)
    ap.add_argument("--no-backward", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    b, n, m = args.b, args.n, args.m
    a = torch.from_numpy(_data.cloud(args.kind, b, n, 1)).to(dev)
    c = torch.from_numpy(_data.cloud(args.kind2 or args.kind, b, m, 2)).to(dev)
    g1, g2 = torch.rand(b, n, device=dev), torch.rand(b, m, device=dev)
    d1, d2 = torch.empty(b, n, device=dev), torch.empty(b, m, device=dev)
    i1, i2 = torch.empty(b, n, device=dev, dtype=torch.int32), torch.empty(b, m, device=dev, dtype=torch.int32)
    gx1, gx2 = torch.empty(b, n, 3, device=dev), torch.empty(b, m, 3, device=dev)
    ws = L.workspace(L.lib.mvp_chamfer_forward_workspace_bytes(b, n, m), dev)
    s = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    P = L.ptr
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for k in range(args.steps):
        e0.record()
        L.check(L.lib.mvp_chamfer_forward_algo(ALGO[args.algo], b, n, m, P(a), P(c), P(d1), P(d2), P(i1), P(i2), P(ws),
                                               ws.numel(), s), "forward")
        e1.record()
        if not args.no_backward:
            L.check(L.lib.mvp_chamfer_backward(b, n, m, P(a), P(c), P(g1), P(g2), P(i1), P(i2), P(gx1), P(gx2), s), "backward")
        e2.record()
        torch.cuda.synchronize()
        print(f"step {k}: forward {e0.elapsed_time(e1) * 1e3:.1f} us, backward {e1.elapsed_time(e2) * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
