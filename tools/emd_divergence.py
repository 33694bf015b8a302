#!/usr/bin/env python
"""Debug aid (GPU box): finds the first auction round at which the reference EMD kernels (oracle/_ref) and
the CPU oracle disagree on a golden input, and prints the GetMax near-tie that caused it.
    gpurun -- 'python tools/emd_divergence.py emd_2048'
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import ref_cuda  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "ref_cuda_golden.npz"))
case = sys.argv[1] if len(sys.argv) > 1 else "emd_2048"
x1, x2 = G[case + ".xyz1"], G[case + ".xyz2"]
eps, iters = float(G[case + ".eps"]), int(G[case + ".iters"])
dev = torch.device("cuda:0")
for cloud in range(x1.shape[0]):
    a, c = x1[cloud], x2[cloud]
    ta, tc = torch.from_numpy(a[None]).to(dev), torch.from_numpy(c[None]).to(dev)
    for policy in (0, 1):
        oracle.emd_set_tie_policy(policy)
        first = None
        for k in range(1, iters + 1):
            rs = {n: v.cpu().numpy().reshape(-1) for n, v in ref_cuda.emd_forward(ta, tc, eps, k, return_state=True).items()}
            os_ = oracle.emd_forward_state(a, c, eps, k)
            if not (rs["assignment"] == os_["assignment"]).all():
                first = k
                break
        print(f"{case} cloud {cloud} policy {policy}: first differing round count = {first}")
        if first is None:
            continue
        # The race happened in round r = first-1 (with iters == r that round is the assign-everyone round, so
        # the assignments still agree, but GetMax ran and left its winners in max_idx).
        r = first - 1
        if r < 1:
            continue
        rs = {n: v.cpu().numpy().reshape(-1) for n, v in ref_cuda.emd_forward(ta, tc, eps, r, return_state=True).items()}
        os_ = oracle.emd_forward_state(a, c, eps, r)
        bid, inc = os_["bid"], os_["bid_increments"]
        print("  round", r, "bid arrays equal:", (rs["bid"] == bid).all(), (rs["bid_increments"] == inc).all())
        un = os_["last_unassigned"]
        dj = np.where(rs["bid_increments"] != inc)[0]
        bc = len(a) // 1024
        upb = (len(un) + bc - 1) // bc
        print(f"  unassigned entering round {r}: {len(un)}; unass_per_block {upb}, thread_per_unass {1024 // max(upb, 1)}")
        print("  sources whose increment differs:", dj[:10], "of which bidding this round:", np.isin(dj, un)[:10])
        for j in dj[:6]:
            print(f"   source {j}: bid {bid[j]} (ref {rs['bid'][j]}), inc oracle {float(inc[j])!r} ref {float(rs['bid_increments'][j])!r}")
        diff_t = np.where(rs["max_idx"] != os_["max_idx"])[0]
        print("  targets whose max_idx differs:", diff_t[:10], "ref", rs["max_idx"][diff_t[:10]], "oracle", os_["max_idx"][diff_t[:10]])
        for t in diff_t[:5]:
            bidders = np.where(bid == t)[0]
            top = float(inc[bidders].max()) if len(bidders) else 0.0
            print(f"   target {t}: bidders near the max (idx, inc):",
                  [(int(j), float(inc[j])) for j in bidders if abs(float(inc[j]) - top) < 1e-5])
oracle.emd_set_tie_policy(0)
