#!/usr/bin/env python
"""Tuning aid: builds libmvp_ops variants of the Chamfer kernels (macro knobs in csrc/chamfer_fused.cu and
csrc/chamfer_grid.cu) into gpurun_build/ (here, no GPU needed) and, on the GPU box, times mvp_chamfer_forward of
each at B=32 N=M=16384 (and at the VRCNet size B=64 N=M=2048 for the grid set).

    python tools/pair_variants.py build [pair|grid]         # in the build container
    gpurun -- 'python tools/pair_variants.py run [pair|grid]'
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_build")
CSRC = os.path.join(ROOT, "mvp_benchmark_b200", "csrc")
GRID_VARIANTS = {
    "base": [],
    "minb10": ["-DMVP_GRID_QMINB=10"],
    "minb12": ["-DMVP_GRID_QMINB=12"],
    "minb16": ["-DMVP_GRID_QMINB=16"],
    "q256minb6": ["-DMVP_GRID_QTHREADS=256", "-DMVP_GRID_QMINB=6"],
    "zbase": [],  # the base library once more, last: the first variant of a run pays the clock ramp
    "q256": ["-DMVP_GRID_QTHREADS=256"],
    "q64": ["-DMVP_GRID_QTHREADS=64"],
    "ppc1": ["-DMVP_GRID_PPC=1"],
    "ppc3": ["-DMVP_GRID_PPC=3"],
    "ppc4": ["-DMVP_GRID_PPC=4"],
    "noxclip": ["-DMVP_GRID_NOXCLIP"],
    "unroll2": ["-DMVP_GRID_SCAN_UNROLL=2"],
    "unroll4": ["-DMVP_GRID_SCAN_UNROLL=4"],
}
EMD_VARIANTS = {
    "base": [],
    "fs0": ["-DMVP_EMD_FULLSCAN_EVALS=0"],
    "fs128": ["-DMVP_EMD_FULLSCAN_EVALS=128"],
    "ppc2": ["-DMVP_EMD_GRID_PPC=2"],
    "ppc8": ["-DMVP_EMD_GRID_PPC=8"],
}
FPS_VARIANTS = {"base": [], "pmax8": ["-DMVP_FPS_PMAX=8"], "pmax4": ["-DMVP_FPS_PMAX=4"]}
SETS = {}
VARIANTS = {
    "base": [],
    "minb3": ["-DMVP_PAIR_MINB=3"],
    "cols2": ["-DMVP_PAIR_COLS=2"],
    "cols2_minb3": ["-DMVP_PAIR_COLS=2", "-DMVP_PAIR_MINB=3"],
    "rows4_minb4": ["-DMVP_PAIR_ROWS=4", "-DMVP_PAIR_MINB=4"],
    "rows4_minb3": ["-DMVP_PAIR_ROWS=4", "-DMVP_PAIR_MINB=3"],
    "scalar": ["-DMVP_PAIR_SCALAR=1"],
    "scalar_minb3": ["-DMVP_PAIR_SCALAR=1", "-DMVP_PAIR_MINB=3"],
}


SETS.update({"pair": VARIANTS, "grid": GRID_VARIANTS, "emd": EMD_VARIANTS, "fps": FPS_VARIANTS})
SYMBOL = {"pair": "chamfer_pair_kernelILi8", "grid": "chamfer_grid_query_kernel", "emd": "emd_auction_grid_kernel",
          "fps": "fps_kernelILi256"}
SHAPES = {"pair": [(32, 16384, 16384)], "grid": [(32, 16384, 16384), (64, 2048, 2048)]}


def build(which):
    os.makedirs(OUT, exist_ok=True)
    for name, flags in SETS[which].items():
        lib = os.path.join(OUT, f"libvar_{which}_{name}.so")
        cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler",
               "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-shared", "-Xptxas", "-v"] + flags + \
              [os.path.join(CSRC, f) for f in (("capi.cu", "emd.cu") if which == "emd" else ("capi.cu", "fps.cu") if which == "fps" else
                                               ("capi.cu", "chamfer.cu", "chamfer_fused.cu", "chamfer_grid.cu", "pointnet2.cu", "pointnet2_staged.cu"))] + ["-o", lib]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            print(name, "FAILED\n", r.stderr[-2000:])
            continue
        lines = (r.stdout + r.stderr).splitlines()
        for i, l in enumerate(lines):
            if SYMBOL[which] in l and "Compiling" in l:
                print(name, "|", lines[i + 1].strip(), "|", lines[i + 2].strip())


def run_emd():
    """mvp_emd_forward of every variant at (b, n, iters) points that stress the first rounds, the tail, and both."""
    import torch
    dev = torch.device("cuda:0")
    ref = {}
    for b, n, iters in [(64, 8192, 50), (32, 2048, 50), (16, 8192, 3000), (32, 2048, 3000), (64, 8192, 5)]:
        print("shape", (b, n, iters), flush=True)
        g = torch.Generator(device=dev)
        g.manual_seed(0)
        x1, x2 = torch.rand(b, n, 3, device=dev, generator=g), torch.rand(b, n, 3, device=dev, generator=g)
        for name in SETS["emd"]:
            path = os.path.join(OUT, f"libvar_emd_{name}.so")
            if not os.path.isfile(path):
                continue
            L = ctypes.CDLL(path)
            L.mvp_emd_forward_workspace_bytes.restype = ctypes.c_size_t
            ws = torch.empty(L.mvp_emd_forward_workspace_bytes(b, n), dtype=torch.uint8, device=dev)
            d, a = torch.empty(b, n, device=dev), torch.empty(b, n, device=dev, dtype=torch.int32)
            P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

            def call():
                rc = L.mvp_emd_forward(b, n, n, P(x1), P(x2), ctypes.c_float(0.005), iters, P(d), P(a), P(ws),
                                       ctypes.c_size_t(ws.numel()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                assert rc == 0, rc
            call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                call()
            e1.record()
            e1.synchronize()
            key = (b, n, iters)
            if key not in ref:
                ref[key] = a.clone()
            print(f"{name:10s} {e0.elapsed_time(e1) / 3:9.3f} ms  same_as_base={torch.equal(a, ref[key])}", flush=True)


def run_fps():
    import torch
    dev = torch.device("cuda:0")
    ref = {}
    for b, n, m in [(32, 2048, 2048), (64, 3072, 2048), (64, 3072, 1536), (64, 1536, 768), (64, 768, 384), (8, 8192, 1024),
                    (4, 16384, 512)]:
        print("shape", (b, n, m), flush=True)
        g = torch.Generator(device=dev)
        g.manual_seed(0)
        x = torch.rand(b, n, 3, device=dev, generator=g)
        for name in SETS["fps"]:
            path = os.path.join(OUT, f"libvar_fps_{name}.so")
            if not os.path.isfile(path):
                continue
            L = ctypes.CDLL(path)
            idx = torch.empty(b, m, device=dev, dtype=torch.int32)
            P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

            def call():
                rc = L.mvp_furthest_point_sampling(b, n, m, P(x), ctypes.c_void_p(0), P(idx),
                                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                assert rc == 0, rc
            call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                call()
            e1.record()
            e1.synchronize()
            key = (b, n, m)
            if key not in ref:
                ref[key] = idx.clone()
            print(f"{name:10s} {e0.elapsed_time(e1) / 5:9.4f} ms  {e0.elapsed_time(e1) / 5 / m * 1e6:7.1f} ns/pick  "
                  f"same_as_base={torch.equal(idx, ref[key])}", flush=True)


def run(which):
    if which == "emd":
        return run_emd()
    if which == "fps":
        return run_fps()
    for shape in SHAPES[which]:
        print("shape", shape, flush=True)
        run_shape(which, *shape)


def run_shape(which, b, n, m):
    import torch
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    x1, x2 = torch.rand(b, n, 3, device=dev, generator=g), torch.rand(b, m, 3, device=dev, generator=g)
    ref = None
    for name in SETS[which]:
        path = os.path.join(OUT, f"libvar_{which}_{name}.so")
        if not os.path.isfile(path):
            continue
        L = ctypes.CDLL(path)
        L.mvp_chamfer_forward_workspace_bytes.restype = ctypes.c_size_t
        wsb = L.mvp_chamfer_forward_workspace_bytes(b, n, m)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        d1, d2 = torch.empty(b, n, device=dev), torch.empty(b, m, device=dev)
        i1, i2 = torch.empty(b, n, device=dev, dtype=torch.int32), torch.empty(b, m, device=dev, dtype=torch.int32)
        P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

        def call():
            rc = L.mvp_chamfer_forward(b, n, m, P(x1), P(x2), P(d1), P(d2), P(i1), P(i2), P(ws), ctypes.c_size_t(wsb),
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, rc

        for _ in range(3):
            call()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        out = (d1.clone(), d2.clone(), i1.clone(), i2.clone())
        if ref is None:
            ref = out
        same = all(torch.equal(a, c) for a, c in zip(out, ref))
        print(f"{name:16s} fwd median {ts[5]:.3f} ms  best {ts[0]:.3f} ms  same_as_base={same}", flush=True)


if __name__ == "__main__":
    which = sys.argv[2] if len(sys.argv) > 2 else "pair"
    build(which) if sys.argv[1] == "build" else run(which)
