#!/usr/bin/env python
"""Generates tests/golden/knn_model_utils.npz by importing the REFERENCE's completion/model_utils.py from
/root/reference and running its pure-torch `knn` (model_utils.py:242-247) and `knn_point` (:250-259) on CPU
tensors — the functions mvp_knn_points / model_patches replace.  Stored per case: the
inputs, the reference's indices, its (negative squared) top-k distances, and the (k+1)-th distance so that the
test can tell a real disagreement from a near-tie the matmul expansion orders differently.
Run in the build container (the reference tree does not travel to the GPU box):

    python tests/golden/make_golden_knn.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MVP_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
import mvp_benchmark_b200  # noqa: E402

mvp_benchmark_b200.install()          # model_utils imports `metrics` / `mm3d_pn2` at module level
sys.path.insert(1, os.path.join(REF, "completion"))
import model_utils  # noqa: E402  (the reference's module)

out = {}
# VRCNet's shapes scaled to CPU size: self-kNN with k=16 (vrcnet.py:246-276), knn_point with pk=10 from a
# subsample to the full cloud (model_utils.py:98), ECG's knn_point(2, ...) on small groups (model_utils.py:217)
for name, (b, n, m, k, seed) in {"self16": (2, 768, 768, 16, 0), "self20": (2, 384, 384, 20, 1),
                                 "point10": (2, 768, 384, 10, 2), "point2_small": (8, 12, 12, 2, 3),
                                 "point4_sphere": (2, 1024, 512, 4, 4)}.items():
    torch.manual_seed(seed)
    cloud = torch.rand(b, n, 3)
    if name.endswith("sphere"):
        cloud = torch.nn.functional.normalize(torch.randn(b, n, 3), dim=2) * 0.5 + 0.5
    if name.startswith("self"):
        queries = cloud
        idx = model_utils.knn(cloud.transpose(1, 2).contiguous(), k)
        full = -(-(cloud ** 2).sum(2, keepdim=True) + 2 * cloud @ cloud.transpose(1, 2)
                 - (cloud ** 2).sum(2).unsqueeze(1))
    else:
        queries = cloud[:, torch.randperm(n)[:m]].contiguous()
        neg, idx = model_utils.knn_point(k, cloud, queries)
        out[name + "_negdist"] = neg.numpy()
        full = -(-(queries ** 2).sum(2, keepdim=True) + 2 * queries @ cloud.transpose(1, 2)
                 - (cloud ** 2).sum(2).unsqueeze(1))
    kth = torch.sort(full, dim=2)[0][:, :, k - 1:k + 1] if n > k else torch.sort(full, dim=2)[0][:, :, k - 1:k].repeat(1, 1, 2)
    out[name + "_cloud"] = cloud.numpy()
    out[name + "_queries"] = queries.numpy()
    out[name + "_idx"] = idx.numpy().astype(np.int32)
    out[name + "_kth_gap"] = (kth[:, :, 1] - kth[:, :, 0]).numpy()      # expansion-distance gap k-th -> (k+1)-th
    out[name + "_k"] = np.int32(k)
np.savez_compressed(os.path.join(HERE, "knn_model_utils.npz"), **out)
print("wrote knn_model_utils.npz", {k: getattr(v, "shape", v) for k, v in out.items()})
