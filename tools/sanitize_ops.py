#!/usr/bin/env python
"""Target for compute-sanitizer: every entry point of libmvp_ops.so once, at sizes that reach each kernel variant
(grid / brute / cluster build / vector backward / sorted FPS / top-k lists / staged and CSR scatters / EMD cluster).

    compute-sanitizer --tool memcheck  python tools/sanitize_ops.py
    compute-sanitizer --tool racecheck python tools/sanitize_ops.py
    compute-sanitizer --tool synccheck python tools/sanitize_ops.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvp_benchmark_b200
mvp_benchmark_b200.install()
import metrics
import mm3d_pn2 as mm
from mvp_benchmark_b200 import fused, _lib

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
R = lambda *s: torch.rand(*s, device=dev, generator=g)  # noqa: E731
small = "--small" in sys.argv


def cd(b, n, m):
    a, c = R(b, n, 3).requires_grad_(True), R(b, m, 3).requires_grad_(True)
    d1, d2, _, _ = metrics.cd()(a, c)
    (d1.sum() + d2.sum()).backward()


cd(2, 300, 200)                    # generic kernels
cd(2, 2048, 1536)                  # grid path, one-CTA build
cd(2, 8192, 7001)                  # grid path, two-CTA cluster build
if not small:
    cd(16, 16384, 16384)           # four-points-per-thread backward
x1 = R(2, 2048, 3)
L = _lib.lib
ws = _lib.workspace(L.mvp_chamfer_forward_workspace_bytes(2, 2048, 2048), dev)
o = [torch.empty(2, 2048, device=dev) for _ in range(2)] + [torch.empty(2, 2048, device=dev, dtype=torch.int32) for _ in range(2)]
x2 = R(2, 2048, 3) + 5.0           # disjoint clouds: hand-over to the fused brute-force kernels
_lib.check(L.mvp_chamfer_forward(2, 2048, 2048, _lib.ptr(x1), _lib.ptr(x2), *[_lib.ptr(t) for t in o], _lib.ptr(ws),
                                 ws.numel(), _lib.stream_of(x1)), "chamfer hand-over")
e1, e2 = R(2, 1024, 3).requires_grad_(True), R(2, 1024, 3)
d, a = metrics.emd()(e1, e2, 0.005, 10)
d.sum().backward()
metrics.emd()(R(2, 4096, 3), R(2, 4096, 3), 0.005, 5)   # cluster of CTAs per cloud
for n, m in ((2048, 512), (5000, 300), (300, 300)):      # unsorted, sorted, tiny
    mm.furthest_point_sample(R(2, n, 3), m)
dm = R(2, 256, 256)
mm.furthest_point_sample_with_dist(dm, 64)
xyz, ctr = R(2, 2048, 3), R(2, 102, 3)
mm.ball_query(0.0, 0.2, 16, xyz, ctr)
mm.knn(8, xyz, ctr)
for k, n, m in ((10, 700, 1536), (16, 1536, 1536), (20, 300, 900), (32, 512, 2048), (40, 100, 500), (3, 50, 60)):
    fused.knn_points(k, R(2, m, 3), R(2, n, 3))
u, kn = R(2, 3072, 3), R(2, 1536, 3)
dist, idx = mm.three_nn(u, kn)
feat = torch.randn(2, 64, 1536, device=dev, generator=g).requires_grad_(True)
w = torch.softmax(-dist, 2).contiguous()
mm.three_interpolate(feat, idx, w).sum().backward()
f2 = torch.randn(2, 32, 3072, device=dev, generator=g).requires_grad_(True)
gi = torch.randint(0, 3072, (2, 15360), device=dev, generator=g, dtype=torch.int32)
mm.gather_points(f2, gi).sum().backward()
gi3 = torch.randint(0, 3072, (2, 1536, 4), device=dev, generator=g, dtype=torch.int32)
f3 = torch.randn(2, 16, 3072, device=dev, generator=g).requires_grad_(True)
mm.grouping_operation(f3, gi3).sum().backward()
la, lc = R(3, 2048).requires_grad_(True), R(3, 1024).requires_grad_(True)
cp, ct = fused.chamfer_loss(la, lc)
(cp.sum() + ct.sum()).backward()
fused.three_nn_weights(u, kn)
# ---- round 2: the completion pass of the Chamfer grid path on the geometries that reach each of its branches
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import _data  # noqa: E402
T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
cdf = metrics.cd()
cdf(T(_data.uniform(2, 4096, 1)), T(_data.blob(2, 4096, 2)))          # far queries, a warp per query
cdf(T(_data.uniform(2, 4096, 1)), T(_data.shifted(2, 4096, 2)))       # disjoint: lanes over queries, list regrouping
cdf(T(_data.clustered(2, 4096, 1)), T(_data.clustered(2, 4096, 2)))   # dense cells: Z-order re-sort (bitonic), plan list B
cdf(T(_data.outliers(2, 2048, 1)), T(_data.outliers(2, 3000, 2)))
cdf(T(_data.constant(2, 2048, 1)), T(_data.uniform(2, 2048, 2)))      # one point n times
if not small:
    cdf(T(_data.uniform(1, 40000, 1)), T(_data.blob(1, 40000, 2)))    # two level-2 nodes
# ---- round 2: fused chains, cluster FPS, warp-per-centre knn, row-wise top-k
p = R(3, 2048, 3).requires_grad_(True)
i_, o_ = fused.fps_gather(p, 512)
o_.sum().backward()
fused.fps_gather(R(2, 6000, 3), 200, channels_first=True)              # four-CTA cluster + epilogue
mm.furthest_point_sample(R(2, 8192, 3), 300)
q = R(2, 2048, 3).requires_grad_(True)
_, gq = fused.ball_query_group(0, 0.1, 12, q, q[:, :102].detach().contiguous())
gq.sum().backward()
mm.knn(16, R(2, 2048, 3), R(2, 300, 3))
mm.knn(100, R(1, 1500, 3), R(1, 77, 3))
mm.knn(9, R(2, 5, 3), R(2, 7, 3))
fused.topk_rows(torch.randn(3, 700, 700, device=dev, generator=g), 16)
fused.topk_rows(torch.randn(5, 33, device=dev, generator=g), 32)
ff = torch.randn(2, 19, 1536, device=dev, generator=g).requires_grad_(True)
fused.gather_max(ff, torch.randint(0, 1536, (2, 700, 10), device=dev, generator=g, dtype=torch.int32)).sum().backward()
yy = torch.randn(2, 16, 768, device=dev, generator=g).requires_grad_(True)
ww = torch.randn(2, 2, 10, 768, device=dev, generator=g).requires_grad_(True)
fused.neighbor_weighted_sum(yy, torch.randint(0, 768, (2, 768, 10), device=dev, generator=g, dtype=torch.int32), ww).sum().backward()
# the 1x1 layers (csrc/pointwise.cu): resident-weight and streaming tcgen05 kernels, masked input gradient, weight + bias
# gradient, bias kernels; sizes off every multiple
for (b_, c_, o_, n_) in ((2, 64, 256, 300), (2, 67, 20, 130), (1, 600, 300, 77), (2, 3, 5, 129)):
    px = torch.randn(b_, c_, n_, device=dev, generator=g).requires_grad_(True)
    pw = torch.randn(o_, c_, 1, device=dev, generator=g).requires_grad_(True)
    pb = torch.randn(o_, device=dev, generator=g).requires_grad_(True)
    fused.pointwise_conv(px, pw, pb, relu=True).sum().backward()
    fused._pointwise_wgrad_raw(torch.randn(b_, o_, n_, device=dev, generator=g), px.detach(), c_ < 256) if c_ <= 256 else None
mxl = torch.randn(2, 5, 77, 20, device=dev, generator=g).requires_grad_(True)
fused.max_last(mxl)[0].sum().backward()
fused.max_last(torch.randn(3, 50, 7, device=dev, generator=g))
cx = torch.randn(2, 40, 515, device=dev, generator=g).requires_grad_(True)
cw = torch.randn(24, 40, 1, device=dev, generator=g).requires_grad_(True)
cb = torch.randn(24, device=dev, generator=g).requires_grad_(True)
fused.conv_bias(cx, cw, cb).sum().backward()
torch.cuda.synchronize()
print("sanitize_ops: all entry points ran,", _lib.launch_count(), "kernels launched")
