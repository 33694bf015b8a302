"""GPU (needs two devices; skipped otherwise): the operators driven the way the reference drives them — ONE process,
several devices, a worker thread per device (nn.DataParallel, completion/train.py:49).  Kernels that opt in to more
than 48 KB of dynamic shared memory must do so on every device they run on, not only on the first one."""
import threading

import numpy as np
import pytest
import torch

import _data

pytestmark = pytest.mark.gpu


def _work(ops, dev, out):
    metrics, mm = ops
    from mvp_benchmark_b200 import fused
    torch.cuda.set_device(dev)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    x1, x2 = T(_data.uniform(2, 8192, 1)), T(_data.uniform(2, 8192, 2))          # cluster build: 96 KB
    a, c = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    d1, d2, i1, i2 = metrics.cd()(a, c)
    (d1.sum() + d2.sum()).backward()
    far = metrics.cd()(x1, x2 + 5.0)                                            # hand-over: fused brute-force kernels
    fps_small = mm.furthest_point_sample(x1[:, :3072].contiguous(), 256)         # cloud staged in shared memory: 48 KB
    fps_big = mm.furthest_point_sample(x1[:, :6000].contiguous(), 128)           # sorted kernel: 120 KB
    feat = torch.randn(2, 16, 8192, device=dev, generator=torch.Generator(device=dev).manual_seed(3), requires_grad=True)
    idx = torch.randint(0, 8192, (2, 16384), device=dev, generator=torch.Generator(device=dev).manual_seed(4), dtype=torch.int32)
    g = mm.gather_points(feat, idx)                                              # staged rows: 64 KB
    g.sum().backward()
    ed, ea = metrics.emd()(x1[:, :4096].contiguous(), x2[:, :4096].contiguous(), 0.005, 5)
    kd, ki = fused.knn_points(16, x1[:, :3072].contiguous())
    torch.cuda.synchronize(dev)
    out[dev.index] = [t.detach().cpu() for t in (d1, d2, i1, i2, a.grad, c.grad, far[0], far[2], fps_small, fps_big, g,
                                                 feat.grad, ed, ea, kd, ki)]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices in one process")
@pytest.mark.parametrize("threaded", [False, True])
def test_one_process_two_devices(ops, threaded):
    devs = [torch.device("cuda", 0), torch.device("cuda", 1)]
    out = {}
    if threaded:
        ts = [threading.Thread(target=_work, args=(ops, d, out)) for d in devs]
        [t.start() for t in ts]
        [t.join() for t in ts]
    else:
        for d in devs:
            _work(ops, d, out)
    assert set(out) == {0, 1}
    names = "d1 d2 i1 i2 gx1 gx2 far_d far_i fps_small fps_big gather gather_grad emd_d emd_a knn_d knn_i".split()
    for nm, p, q in zip(names, out[0], out[1]):
        if nm in ("gx1", "gx2", "gather_grad"):
            torch.testing.assert_close(p, q, rtol=1e-5, atol=1e-6, msg=nm)      # atomically accumulated
        else:
            assert torch.equal(p, q), nm
