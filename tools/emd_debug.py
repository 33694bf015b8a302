import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _data
from _impls import CudaImpl
g = CudaImpl("cuda:0")
for n, b in [(1024, 1), (2048, 2), (3072, 1)]:
    x1, x2 = _data.uniform(b, n, 31), _data.uniform(b, n, 32)
    for iters in (1, 2, 3, 5, 10):
        wd, wa = g.emd_forward(x1, x2, 0.005, iters, algo="brute")
        d, a = g.emd_forward(x1, x2, 0.005, iters, algo="grid")
        bad = np.argwhere(a != wa)
        print(n, b, "iters", iters, "mismatch", len(bad))
        for (bb, j) in bad[:3]:
            p = x1[bb, j]
            dg = np.linalg.norm(x2[bb, a[bb, j]] - p); dw = np.linalg.norm(x2[bb, wa[bb, j]] - p)
            print("   src", bb, j, "grid->", a[bb, j], dg, "brute->", wa[bb, j], dw)
