#!/usr/bin/env python
"""Tuning aid: builds libmvp_ops variants of the Chamfer pair kernel (macro knobs in csrc/chamfer_fused.cu) into
gpurun_build/ (here, no GPU needed) and, on the GPU box, times mvp_chamfer_forward of each at B=32 N=M=16384.

    python tools/pair_variants.py build          # in the build container
    gpurun -- 'python tools/pair_variants.py run'
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_build")
CSRC = os.path.join(ROOT, "mvp_benchmark_b200", "csrc")
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


def build():
    os.makedirs(OUT, exist_ok=True)
    for name, flags in VARIANTS.items():
        lib = os.path.join(OUT, f"libvar_{name}.so")
        cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler",
               "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-shared", "-Xptxas", "-v"] + flags + \
              [os.path.join(CSRC, f) for f in ("capi.cu", "chamfer.cu", "chamfer_fused.cu")] + ["-o", lib]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            print(name, "FAILED\n", r.stderr[-2000:])
            continue
        lines = (r.stdout + r.stderr).splitlines()
        for i, l in enumerate(lines):
            if "chamfer_pair_kernelILi8" in l:
                print(name, "|", lines[i + 1].strip(), "|", lines[i + 2].strip())


def run():
    import torch
    dev = torch.device("cuda:0")
    b, n, m = 32, 16384, 16384
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    x1, x2 = torch.rand(b, n, 3, device=dev, generator=g), torch.rand(b, m, 3, device=dev, generator=g)
    ref = None
    for name in VARIANTS:
        path = os.path.join(OUT, f"libvar_{name}.so")
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
    build() if sys.argv[1:] == ["build"] else run()
