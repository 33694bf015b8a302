#!/usr/bin/env python
"""BASELINE config C5: the auction EMD standalone, B = 64 clouds of n = 8192 points per GPU, eps 0.005, 50 rounds —
weak scaling over the ranks of a torchrun launch (no collective on the operator path: every cloud is independent).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/emd_scale.py
"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvp_benchmark_b200 import dist as mdist

rank, world, local = mdist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
import mvp_benchmark_b200  # noqa: E402
mvp_benchmark_b200.install()
import metrics  # noqa: E402

b, n, eps, iters, steps = 64, 8192, 0.005, 50, 10
g = torch.Generator(device=dev).manual_seed(100 + rank)
x1, x2 = torch.rand(b, n, 3, device=dev, generator=g), torch.rand(b, n, 3, device=dev, generator=g)
emd = metrics.emd()
for _ in range(3):
    emd(x1, x2, eps, iters)
torch.cuda.synchronize()
mdist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    d, a = emd(x1, x2, eps, iters)
e1.record()
torch.cuda.synchronize()
mdist.barrier()
ms = mdist.max_over_ranks(e0.elapsed_time(e1), dev) / steps
if rank == 0:
    print(json.dumps({"workload": f"EMD auction B={b} n={n} eps={eps} iters={iters} per GPU", "n_gpus": world, "ms_per_call": ms,
                      "clouds_per_s": world * b / (ms * 1e-3), "point_pairs_per_s": world * b * float(n) * n / (ms * 1e-3),
                      "mean_sqrt_dist": float(torch.sqrt(d).mean())}))
