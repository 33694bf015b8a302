#!/usr/bin/env python
"""Probe: host->device bandwidth of several pinned buffers allocated one after the other in one process (is a slow
host->device copy a property of the allocation or of the box?), repeated over time."""
import time
import torch
dev = torch.device("cuda:0")
d = torch.empty(6291456, dtype=torch.uint8, device=dev)


def bw(h, iters=10):
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return h.numel() * iters / e0.elapsed_time(e1) / 1e6


bufs = []
for i in range(8):
    h = torch.empty(6291456, dtype=torch.uint8).pin_memory()
    h.fill_(i)
    bufs.append(h)
for rnd in range(3):
    print("round", rnd, " ".join(f"{bw(h):6.1f}" for h in bufs), "GB/s", flush=True)
    time.sleep(0.5)
src = torch.rand(32, 16384, 3, device=dev)
hs = [src.cpu().pin_memory() for _ in range(4)]
print("x.cpu().pin_memory():", " ".join(f"{bw(h.view(torch.uint8).view(-1)):6.1f}" for h in hs), "GB/s")
