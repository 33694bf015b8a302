#!/usr/bin/env python
"""Probe: does host->device bandwidth from pinned memory depend on where the pages live?  Prints the NUMA layout the
container sees, the GPU's ideal CPU set (NVML), and H2D / D2H bandwidth for pinned buffers allocated before and after
binding the process to that CPU set."""
import glob
import os
import torch

dev = torch.device("cuda:0")
print("cpus allowed:", sorted(os.sched_getaffinity(0)))
for node in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    try:
        print(os.path.basename(node), "cpulist", open(node + "/cpulist").read().strip())
    except OSError:
        pass
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    ideal = [64 * w + bit for w, word in enumerate(mask) for bit in range(64) if (word >> bit) & 1]
    print("GPU 0 ideal cpus (NVML):", ideal[:8], "...", len(ideal))
    try:
        print("GPU 0 numa node:", open(f"/sys/bus/pci/devices/{pynvml.nvmlDeviceGetPciInfo(h).busId.lower()[4:]}/numa_node").read().strip())
    except Exception as e:
        print("numa_node unreadable:", e)
except Exception as e:
    ideal = []
    print("NVML:", e)


def bw(tag):
    hin = torch.empty(12582912, dtype=torch.uint8).pin_memory()
    hin.fill_(1)
    hout = torch.empty(20971520, dtype=torch.uint8).pin_memory()
    din = torch.empty(12582912, dtype=torch.uint8, device=dev)
    dout = torch.empty(20971520, dtype=torch.uint8, device=dev)
    for name, fn, nb in (("H2D", lambda: din.copy_(hin, non_blocking=True), 12582912),
                         ("D2H", lambda: hout.copy_(dout, non_blocking=True), 20971520)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{tag:28s} {name} {ms:7.4f} ms {nb / ms / 1e6:6.1f} GB/s")


bw("default placement")
if ideal:
    allowed = set(os.sched_getaffinity(0)) & set(ideal)
    if allowed:
        os.sched_setaffinity(0, allowed)
        print("bound to", len(allowed), "cpus")
        bw("after binding to ideal cpus")
    else:
        print("none of the ideal cpus is in the allowed set")
for cpu in sorted(os.sched_getaffinity(0))[:1] + sorted(os.sched_getaffinity(0))[-1:]:
    pass
