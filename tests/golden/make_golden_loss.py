#!/usr/bin/env python
"""Generates tests/golden/calc_cd_epilogue.npz by running the REFERENCE's own `calc_cd`
(completion/model_utils.py:67-77) on CPU: `model_utils.cd` is pointed at the reference's pure-torch Chamfer
(utils/metrics/CD/chamfer_python.py:18-39) so that the whole call — distances and the sqrt / mean epilogue — is
reference code.  Stored: dist1, dist2 (the epilogue's inputs) and cd_p, cd_t (its outputs).

    python tests/golden/make_golden_loss.py
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
sys.path.insert(1, os.path.join(REF, "utils", "metrics", "CD"))
import chamfer_python  # noqa: E402  (the reference's modules)
import model_utils  # noqa: E402


class _RefCD(torch.nn.Module):
    def forward(self, a, b):
        return chamfer_python.distChamfer(a, b)


model_utils.cd = _RefCD
out = {}
for name, (b, n, m, seed) in {"vrcnet": (4, 2048, 1024, 0), "square": (3, 1500, 1500, 1), "tiny": (5, 7, 3, 2)}.items():
    torch.manual_seed(seed)
    gt, pred = torch.rand(b, n, 3), torch.rand(b, m, 3)
    dist1, dist2, _, _ = chamfer_python.distChamfer(gt, pred)
    cd_p, cd_t = model_utils.calc_cd(pred, gt)
    out[name + "_dist1"] = dist1.float().numpy()
    out[name + "_dist2"] = dist2.float().numpy()
    out[name + "_cd_p"] = cd_p.float().numpy()
    out[name + "_cd_t"] = cd_t.float().numpy()
np.savez_compressed(os.path.join(HERE, "calc_cd_epilogue.npz"), **out)
print("wrote calc_cd_epilogue.npz", {k: v.shape for k, v in out.items()})
