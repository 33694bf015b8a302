#!/usr/bin/env python
"""EMD parity at BASELINE config C5 (B=64, n=8192, eps 0.005, 50 rounds) — and at the completion models' size
(B=32, n=2048) — against the LIVE reference kernels, cloud by cloud.  Run on a B200:

    python tools/emd_c5_parity.py --json gpurun_out/emd_c5_parity.json

For every cloud: the oracle's count of GetMax windows with two or more in-window bidders (emd_cuda.cu:188-191, a
last-writer race in the reference; `oracle.emd_last_ambiguous_per_cloud`), whether two runs of the reference agree
with each other, whether the product equals reference run A bit for bit (assignment and dist), the fraction of equal
assignments, and the relative error of the mean transport cost mean(sqrt(dist)).  The committed copy of the output
(profiles/r2_emd_c5_parity.json) is where the tolerance of tests/test_gpu_parity.py::test_emd_c5_* comes from.
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
import _impls  # noqa: E402
import oracle  # noqa: E402
from oracle import ref_cuda  # noqa: E402


def one(gpu, dev, kind, b, n, eps, iters, seed):
    x1, x2 = _data.cloud(kind, b, n, seed), _data.cloud(kind, b, n, seed + 1)
    d, a = gpu.emd_forward(x1, x2, eps, iters)
    t1, t2 = torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev)
    runs = []
    for _ in range(2):
        rd, ra = ref_cuda.emd_forward(t1, t2, eps, iters)
        torch.cuda.synchronize()
        runs.append((rd.cpu().numpy(), ra.cpu().numpy()))
    od, oa = oracle.emd_forward(x1, x2, eps, iters)
    amb = oracle.emd_last_ambiguous_per_cloud(b)
    rows = []
    for c in range(b):
        (rdA, raA), (rdB, raB) = (runs[0][0][c], runs[0][1][c]), (runs[1][0][c], runs[1][1][c])
        mean = lambda v: float(np.sqrt(v.astype(np.float64)).mean())
        rows.append({
            "cloud": c, "oracle_race_windows": int(amb[c]),
            "reference_runs_agree": bool((raA == raB).all() and (rdA.view(np.uint32) == rdB.view(np.uint32)).all()),
            "product_equals_reference": bool((a[c] == raA).all() and (d[c].view(np.uint32) == rdA.view(np.uint32)).all()),
            "product_equals_oracle": bool((a[c] == oa[c]).all() and (d[c].view(np.uint32) == od[c].view(np.uint32)).all()),
            "equal_assignments": float((a[c] == raA).mean()),
            "ref_vs_ref_equal_assignments": float((raA == raB).mean()),
            "mean_cost_product": mean(d[c]), "mean_cost_reference": mean(rdA),
            "mean_cost_rel_err": abs(mean(d[c]) - mean(rdA)) / mean(rdA),
            "ref_vs_ref_mean_cost_rel_err": abs(mean(rdB) - mean(rdA)) / mean(rdA),
        })
    free = [r for r in rows if r["oracle_race_windows"] == 0]
    raced = [r for r in rows if r["oracle_race_windows"] > 0]
    summary = {
        "kind": kind, "b": b, "n": n, "eps": eps, "iters": iters, "seed": seed,
        "race_free_clouds": len(free), "raced_clouds": len(raced),
        "race_free_identical": sum(r["product_equals_reference"] for r in free),
        "raced_identical": sum(r["product_equals_reference"] for r in raced),
        "product_equals_oracle_all": all(r["product_equals_oracle"] for r in rows),
        "reference_self_disagreements": sum(not r["reference_runs_agree"] for r in rows),
        "max_mean_cost_rel_err_raced": max([r["mean_cost_rel_err"] for r in raced], default=0.0),
        "max_mean_cost_rel_err_ref_vs_ref": max(r["ref_vs_ref_mean_cost_rel_err"] for r in rows),
        "min_equal_assignments_raced": min([r["equal_assignments"] for r in raced], default=1.0),
        "batch_mean_cost_rel_err": abs(np.sqrt(d.astype(np.float64)).mean() - np.sqrt(runs[0][0].astype(np.float64)).mean())
        / np.sqrt(runs[0][0].astype(np.float64)).mean(),
    }
    print(json.dumps(summary), flush=True)
    return {"summary": summary, "clouds": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--quick", action="store_true", help="C5 at B=8 only")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    gpu = _impls.CudaImpl()
    cases = [("uniform", 64, 8192, 0.005, 50, 41), ("uniform", 64, 8192, 0.005, 50, 141), ("sphere", 64, 8192, 0.005, 50, 43),
             ("uniform", 32, 2048, 0.005, 50, 45), ("uniform", 32, 2048, 0.004, 3000, 47), ("uniform", 8, 8192, 0.002, 3000, 49)]
    if args.quick:
        cases = [("uniform", 8, 8192, 0.005, 50, 41)]
    out = [one(gpu, dev, *c) for c in cases]
    if args.json:
        os.makedirs(os.path.dirname(args.json) or ".", exist_ok=True)
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
