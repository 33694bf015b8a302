import torch, sys
sys.path.insert(0, "/root/repo")
from mvp_benchmark_b200 import _lib
L, P = _lib.lib, _lib.ptr
dev = torch.device("cuda:0")
bb, c, nn, mp = 64, 64, 3072, 15360
go = torch.randn(bb, c, mp, device=dev); idx = torch.randint(0, nn, (bb, mp), device=dev, dtype=torch.int32)
gp = torch.empty(bb, c, nn, device=dev)
ws = _lib.workspace(L.mvp_scatter_workspace_bytes(bb, nn, mp), dev)
S = _lib.stream_of(go)
for _ in range(3):
    L.mvp_gather_points_grad_ws(bb, c, nn, mp, P(go), P(idx), P(gp), P(ws), ws.numel(), S)
    L.mvp_gather_points_grad(bb, c, nn, mp, P(go), P(idx), P(gp), S)
f = torch.randn(bb, c, nn, device=dev); out = torch.empty(bb, c, mp, device=dev)
for _ in range(3):
    L.mvp_gather_points(bb, c, nn, mp, P(f), P(idx), P(out), S)
torch.cuda.synchronize()
