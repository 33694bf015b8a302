#!/usr/bin/env python
"""Probe: device->host bandwidth for the 21 MB of outputs of one Chamfer step (six tensors) from one stream, from two
streams, and as one contiguous buffer; host->device for the 12.6 MB of inputs; both directions at once."""
import torch
dev = torch.device("cuda:0")
sizes = [2097152, 2097152, 2097152, 2097152, 6291456, 6291456]  # bytes: d1 d2 i1 i2 g1 g2 at B=32, N=M=16384
src = [torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes]
dst = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes]
big_src = torch.empty(sum(sizes), dtype=torch.uint8, device=dev)
big_dst = torch.empty(sum(sizes), dtype=torch.uint8).pin_memory()
hin = torch.empty(12582912, dtype=torch.uint8).pin_memory()
din = torch.empty(12582912, dtype=torch.uint8, device=dev)
s1, s2, s3 = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    for s in (s1, s2, s3):
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def one_stream():
    with torch.cuda.stream(s1):
        for a, b in zip(src, dst):
            b.copy_(a, non_blocking=True)


def two_streams():
    for i, (a, b) in enumerate(zip(src, dst)):
        with torch.cuda.stream(s1 if i in (0, 1, 4) else s2):
            b.copy_(a, non_blocking=True)


def one_buffer():
    with torch.cuda.stream(s1):
        big_dst.copy_(big_src, non_blocking=True)


def h2d():
    with torch.cuda.stream(s3):
        din.copy_(hin, non_blocking=True)


def both():
    one_stream(), h2d()


tot = sum(sizes)
for name, fn, nbytes in (("D2H six tensors, one stream", one_stream, tot), ("D2H six tensors, two streams", two_streams, tot),
                         ("D2H one 21 MB buffer", one_buffer, tot), ("H2D 12.6 MB", h2d, 12582912),
                         ("D2H (one stream) + H2D concurrently", both, tot)):
    ms = timed(fn)
    print(f"{name:40s} {ms:7.4f} ms  {nbytes / ms / 1e6:7.1f} GB/s (of the D2H bytes)" if "H2D 12" not in name else
          f"{name:40s} {ms:7.4f} ms  {nbytes / ms / 1e6:7.1f} GB/s")
