#!/usr/bin/env python
"""Times mvp_emd_forward's Bid searches (full scan vs grid-pruned) on a B200 and checks that they agree bit for bit.

    python tools/emd_algos.py [--json gpurun_out/emd_algos.json] [--cases kind:b:n:iters,...]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _data  # noqa: E402
from mvp_benchmark_b200 import _lib as L  # noqa: E402

ALGO = {"brute": 1, "grid": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--cases", default=None)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cases = [("uniform", 64, 8192, 50), ("uniform", 32, 2048, 50), ("sphere", 64, 8192, 50), ("uniform", 16, 8192, 3000),
             ("uniform", 32, 2048, 3000), ("clustered", 32, 2048, 50), ("uniform", 64, 4096, 50)]
    if args.cases:
        cases = [(k, int(b), int(n), int(i)) for k, b, n, i in (c.split(":") for c in args.cases.split(","))]
    out = []
    for kind, b, n, iters in cases:
        a = torch.from_numpy(_data.cloud(kind, b, n, 41)).to(dev)
        c = torch.from_numpy(_data.cloud(kind, b, n, 42)).to(dev)
        ws = L.workspace(L.lib.mvp_emd_forward_workspace_bytes(b, n), dev)
        res = {}
        for algo, code in ALGO.items():
            d = torch.empty(b, n, device=dev)
            asg = torch.empty(b, n, device=dev, dtype=torch.int32)

            def run():
                L.check(L.lib.mvp_emd_forward_algo(code, b, n, n, L.ptr(a), L.ptr(c), 0.005, iters, L.ptr(d), L.ptr(asg),
                                                   L.ptr(ws), ws.numel(), L.stream_of(a)), "emd " + algo)
            run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            res[algo] = (e0.elapsed_time(e1) / args.reps, d.clone(), asg.clone())
        same = torch.equal(res["brute"][2], res["grid"][2]) and torch.equal(res["brute"][1].view(torch.int32),
                                                                            res["grid"][1].view(torch.int32))
        row = {"kind": kind, "b": b, "n": n, "iters": iters, "brute_ms": round(res["brute"][0], 3),
               "grid_ms": round(res["grid"][0], 3), "speedup": round(res["brute"][0] / res["grid"][0], 2),
               "identical": bool(same)}
        print(row, flush=True)
        out.append(row)
    if args.json:
        os.makedirs(os.path.dirname(args.json), exist_ok=True)
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
