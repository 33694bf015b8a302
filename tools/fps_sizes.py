#!/usr/bin/env python
"""Timing aid: mvp_furthest_point_sampling at the VRCNet / ECG shapes (CUDA events, median of 10); run once with
MVP_FPS_SORTED=0 for the unsorted kernel."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvp_benchmark_b200 import _lib

L, P = _lib.lib, _lib.ptr
dev = torch.device("cuda:0")
kinds = {"uniform": lambda b, n: torch.rand(b, n, 3, device=dev),
         "sphere": lambda b, n: torch.nn.functional.normalize(torch.randn(b, n, 3, device=dev), dim=2) * 0.5 + 0.5}
for kind, make in kinds.items():
    for b, n, m in ((32, 2048, 2048), (64, 3072, 2048), (64, 3072, 1536), (64, 1536, 768), (64, 768, 384), (32, 2048, 1024), (16, 8192, 2048)):
        x = make(b, n)
        idx = torch.empty(b, m, device=dev, dtype=torch.int32)
        S = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        call = lambda: _lib.check(L.mvp_furthest_point_sampling(b, n, m, P(x), None, P(idx), S), "fps")  # noqa: E731
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[5]
        print(f"{kind:8s} {b}x{n}->{m}: {t:7.4f} ms  {t / (m - 1) * 1e6:6.1f} ns/pick  checksum {int(idx.long().sum())}", flush=True)
