#!/usr/bin/env python
"""Generates tests/golden/chamfer_python_ref.npz by importing the REFERENCE's own pure-torch Chamfer
(utils/metrics/CD/chamfer_python.py:18-39) from /root/reference and running it on CPU, on the shapes of
its unit test (utils/metrics/CD/unit_test.py:15-16: (4,100,3) vs (4,200,3)) plus BASELINE config C1
((4,2048,3) x2).  Run in the build container (the reference tree does not travel to the GPU box).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("MVP_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(REF, "utils", "metrics", "CD"))
import chamfer_python  # noqa: E402  (the reference's module)

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
for name, (b, n, m, seed) in {"unit": (4, 100, 200, 0), "unit1": (4, 100, 200, 1), "c1": (4, 2048, 2048, 0)}.items():
    torch.manual_seed(seed)
    a = torch.rand(b, n, 3)
    c = torch.rand(b, m, 3)
    d1, d2, i1, i2 = chamfer_python.distChamfer(a, c)
    out[name + "_xyz1"] = a.numpy()
    out[name + "_xyz2"] = c.numpy()
    out[name + "_dist1"] = d1.numpy()
    out[name + "_dist2"] = d2.numpy()
    out[name + "_idx1"] = i1.numpy()
    out[name + "_idx2"] = i2.numpy()
np.savez_compressed(os.path.join(HERE, "chamfer_python_ref.npz"), **out)
print("wrote", os.path.join(HERE, "chamfer_python_ref.npz"), {k: v.shape for k, v in out.items()})
