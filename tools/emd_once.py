#!/usr/bin/env python
"""Runs mvp_emd_forward a few times at (B, n, iters) — a target for `ncu -k regex:emd`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvp_benchmark_b200
mvp_benchmark_b200.install()
import metrics
b, n, iters = (int(v) for v in (sys.argv[1:4] + ["64", "8192", "50"][len(sys.argv) - 1:]))
x1, x2 = torch.rand(b, n, 3, device="cuda"), torch.rand(b, n, 3, device="cuda")
for _ in range(3):
    d, a = metrics.emd()(x1, x2, 0.005, iters)
torch.cuda.synchronize()
print("unique targets in cloud 0:", a[0].unique().numel(), "of", n)
