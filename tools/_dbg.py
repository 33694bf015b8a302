import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _data
from mvp_benchmark_b200 import _lib as L
dev = torch.device("cuda:0")
for kind, b, n, m in [("uniform", 32, 16384, 16384), ("uniform", 64, 2048, 2048), ("sphere", 32, 16384, 16384)]:
    a = torch.from_numpy(_data.cloud(kind, b, n, 1)).to(dev); c = torch.from_numpy(_data.cloud(kind, b, m, 2)).to(dev)
    d1, d2 = torch.empty(b, n, device=dev), torch.empty(b, m, device=dev)
    i1, i2 = torch.empty(b, n, device=dev, dtype=torch.int32), torch.empty(b, m, device=dev, dtype=torch.int32)
    ws = L.workspace(L.lib.mvp_chamfer_forward_workspace_bytes(b, n, m), dev)
    s = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    P = L.ptr
    L.check(L.lib.mvp_chamfer_forward_algo(2, b, n, m, P(a), P(c), P(d1), P(d2), P(i1), P(i2), P(ws), ws.numel(), s), "f")
    torch.cuda.synchronize()
    raw = ws[: b * 64 + 2 * b * 4].cpu().numpy()
    hdr = raw[: b * 64].view(np.int32).reshape(b, 16)
    cnt = raw[b * 64:].view(np.int32)
    print(kind, b, n, m, "leftover total", cnt.sum(), "max", cnt.max(), "first", cnt[:6])
    print(" hdr0 g", hdr[0, 5:8], "ncell", hdr[0, 8], "valid", hdr[0, 9], "blocks", hdr[0, 10:16], "s", hdr[0, 4:5].view(np.float32))
