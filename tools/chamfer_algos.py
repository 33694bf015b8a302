#!/usr/bin/env python
"""Times mvp_chamfer_forward's algorithms (brute force vs grid-pruned) on a B200 at several shapes and point
distributions, checking on the way that they agree bit for bit.  Run under gpurun:

    python tools/chamfer_algos.py [--json gpurun_out/chamfer_algos.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _data  # noqa: E402
from mvp_benchmark_b200 import _lib as L  # noqa: E402

ALGO = {"brute": 1, "grid": 2}


def run(algo, a, c, outs, ws):
    b, n, _ = a.shape
    m = c.shape[1]
    s = L.stream_of(a)
    rc = L.lib.mvp_chamfer_forward_algo(ALGO[algo], b, n, m, L.ptr(a), L.ptr(c), *[L.ptr(o) for o in outs], L.ptr(ws),
                                        ws.numel(), s)
    L.check(rc, "chamfer " + algo)


def time_ms(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--cases", default=None, help="comma-separated kind:b:n:m (default: a built-in sweep)")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cases = [("uniform", 32, 16384, 16384), ("sphere", 32, 16384, 16384), ("clustered", 32, 16384, 16384),
             ("planar", 32, 16384, 16384), ("shifted", 32, 16384, 16384), ("constant", 8, 16384, 16384),
             ("uniform", 64, 2048, 2048), ("uniform", 64, 2048, 3072), ("uniform", 64, 2048, 1024),
             ("sphere", 64, 2048, 2048), ("uniform", 32, 16384, 1024), ("uniform", 4, 2048, 2048),
             ("blob", 32, 16384, 16384), ("blob", 32, 16384, 1024)]
    if args.cases:
        cases = [(k, int(b), int(n), int(m)) for k, b, n, m in (c.split(":") for c in args.cases.split(","))]
    out = []
    for kind, b, n, m in cases:
        # "blob": PCN at initialisation (BASELINE config C2) — the ground truth fills the unit cube, the prediction is a
        # small blob inside it
        a = torch.from_numpy(_data.cloud("uniform" if kind == "blob" else kind, b, n, 1)).to(dev)
        c = torch.from_numpy(_data.cloud(kind, b, m, 2)).to(dev)
        ws = L.workspace(L.lib.mvp_chamfer_forward_workspace_bytes(b, n, m), dev)
        res = {}
        for algo in ALGO:
            outs = [torch.empty(b, n, device=dev), torch.empty(b, m, device=dev),
                    torch.empty(b, n, device=dev, dtype=torch.int32), torch.empty(b, m, device=dev, dtype=torch.int32)]
            ms = time_ms(lambda: run(algo, a, c, outs, ws), reps=args.reps)
            res[algo] = (ms, [o.clone() for o in outs])
        same = all(torch.equal(x.view(torch.int32), y.view(torch.int32)) for x, y in zip(res["brute"][1], res["grid"][1]))
        row = {"kind": kind, "b": b, "n": n, "m": m, "brute_ms": round(res["brute"][0], 4), "grid_ms": round(res["grid"][0], 4),
               "speedup": round(res["brute"][0] / res["grid"][0], 2), "identical": bool(same)}
        print(row, flush=True)
        out.append(row)
    if args.json:
        os.makedirs(os.path.dirname(args.json), exist_ok=True)
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
