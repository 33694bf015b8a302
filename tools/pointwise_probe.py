"""Diagnostic for csrc/pointwise.cu (run under gpurun): one-hot operand-layout checks, random cases against an fp64
matmul, and timings against cuDNN's conv (+ bias) and torch.baddbmm on the completion models' layer shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mvp_benchmark_b200 import fused

dev = torch.device("cuda:0")
torch.manual_seed(0)


def onehot():
    bad = 0
    for (C, O, N, k, co, p) in [(32, 16, 128, 0, 0, 0), (32, 16, 128, 5, 3, 77), (64, 256, 256, 37, 200, 130),
                                (8, 16, 128, 7, 15, 127), (40, 48, 300, 33, 47, 299)]:
        x = torch.zeros(1, C, N, device=dev); x[0, k, p] = 1.0
        w = torch.zeros(O, C, device=dev); w[co, k] = 2.0
        y = fused.pointwise_conv(x, w)
        nz = torch.nonzero(y)
        ok = nz.shape[0] == 1 and nz[0].tolist() == [0, co, p] and float(y[0, co, p]) == 2.0
        bad += not ok
        print("onehot", (C, O, N, k, co, p), "OK" if ok else "BAD: nonzeros %s" % nz[:8].tolist(), flush=True)
    return bad


def rand_cases():
    bad = 0
    for (B, C, O, N, relu, bias) in [(2, 32, 16, 128, False, False), (2, 64, 256, 384, False, True), (3, 3, 5, 77, True, True),
                                     (2, 67, 300, 130, False, True), (2, 512, 1024, 256, True, True), (1, 130, 16, 1000, False, False),
                                     (2, 8, 4, 2048, False, True)]:
        x = torch.randn(B, C, N, device=dev)
        w = torch.randn(O, C, device=dev) / C ** 0.5
        bs = torch.randn(O, device=dev) if bias else None
        y = fused.pointwise_conv(x, w, bs, relu)
        ref = torch.matmul(w.double(), x.double())
        if bias:
            ref = ref + bs.double().view(1, -1, 1)
        if relu:
            ref = ref.clamp_min(0)
        err = float((y.double() - ref).abs().max()); scale = float(ref.abs().max())
        ok = err <= 2e-3 * scale
        bad += not ok
        print("random", (B, C, O, N, relu, bias), "max err %.3e of %.2f" % (err, scale), "OK" if ok else "BAD", flush=True)
    return bad


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def timings():
    out = {}
    for (B, C, O, N) in [(64, 64, 256, 3072), (64, 64, 16, 3072), (64, 128, 256, 2048), (64, 256, 256, 768), (64, 512, 512, 2048),
                         (64, 512, 1024, 2048), (64, 1536, 512, 384), (64, 512, 512, 384), (64, 3, 64, 2048), (64, 256, 128, 1536)]:
        x = torch.randn(B, C, N, device=dev); w = torch.randn(O, C, device=dev) / C ** 0.5; bs = torch.randn(O, device=dev)
        g = torch.randn(B, O, N, device=dev); wt = w.t().contiguous()
        w3 = w.view(O, C, 1)
        row = {"ours_fwd": timeit(lambda: fused._pointwise_conv_raw(x, w, bs)),
               "cudnn_fwd_bias": timeit(lambda: F.conv1d(x, w3, bs)),
               "cudnn_fwd_nobias": timeit(lambda: F.conv1d(x, w3)),
               "baddbmm_fp32": timeit(lambda: torch.baddbmm(bs.view(1, -1, 1), w.view(1, O, C).expand(B, -1, -1), x)),
               "ours_dgrad": timeit(lambda: fused._pointwise_conv_raw(g, wt, None)),
               "cudnn_dgrad": timeit(lambda: torch.ops.aten.convolution_backward(g.unsqueeze(-1), x.unsqueeze(-1), w.view(O, C, 1, 1), None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [True, False, False])),
               "ours_wgrad_bias": timeit(lambda: fused._pointwise_wgrad_raw(g, x, True)) if C < 256 else None,
               "bmm_wgrad": timeit(lambda: torch.bmm(g, x.transpose(1, 2)).sum(0)),
               "cudnn_wgrad": timeit(lambda: torch.ops.aten.convolution_backward(g.unsqueeze(-1), x.unsqueeze(-1), w.view(O, C, 1, 1), None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [False, True, False])),
               "bytes_MB": (B * (C + O) * N * 4) / 1e6, "gflop": 2.0 * B * C * O * N / 1e9}
        out["%dx%d->%dx%d" % (B, C, O, N)] = row
        print("%dx%d->%dx%d" % (B, C, O, N), json.dumps({k: round(v, 4) for k, v in row.items() if v is not None}), flush=True)
    return out


def wgrad_cases():
    bad = 0
    for (B, C, O, N, bias) in [(2, 32, 16, 128, True), (3, 64, 256, 384, True), (3, 3, 5, 77, True), (2, 67, 300, 130, False),
                               (4, 255, 40, 515, True), (2, 256, 128, 1000, False), (64, 64, 4, 3072, True), (1, 8, 8, 4, True)]:
        x = torch.randn(B, C, N, device=dev); g = torch.randn(B, O, N, device=dev)
        gw, gb = fused._pointwise_wgrad_raw(g, x, bias)
        ref = torch.einsum("bon,bcn->oc", g.double(), x.double())
        e1 = float((gw.double() - ref).abs().max() / ref.abs().max())
        e2 = float((gb.double() - g.double().sum((0, 2))).abs().max() / g.double().sum((0, 2)).abs().max()) if bias else 0.0
        ok = e1 < 3e-3 and e2 < 3e-3
        bad += not ok
        print("wgrad", (B, C, O, N, bias), "gw err %.2e gb err %.2e" % (e1, e2), "OK" if ok else "BAD", flush=True)
    return bad


def bias_pieces():
    bad = 0
    for (B, C, N) in [(64, 1024, 2048), (64, 512, 2048), (64, 256, 3072), (3, 5, 77), (64, 4, 3072), (2, 1000, 6)]:
        y = torch.randn(B, C, N, device=dev); bs = torch.randn(C, device=dev)
        ref = y + bs.view(1, -1, 1)
        got = fused.bias_add_(y.clone(), bs)
        ok1 = torch.equal(ref, got)
        s_ref = y.double().sum((0, 2)); s_got = fused.channel_sum(y)
        err = float((s_got.double() - s_ref).abs().max())
        ok2 = err <= 1e-5 * float(y.abs().double().sum((0, 2)).max())
        bad += (not ok1) + (not ok2)
        yy = y.clone()
        row = {"bias_add_ours": timeit(lambda: fused.bias_add_(yy, bs)), "bias_add_torch": timeit(lambda: yy.add_(bs.view(1, -1, 1))),
               "channel_sum_ours": timeit(lambda: fused.channel_sum(y)), "sum_torch": timeit(lambda: y.sum((0, 2))), "MB": B * C * N * 4 / 1e6}
        print("bias", (B, C, N), "add exact" if ok1 else "ADD BAD", "sum err %.2e" % err, "OK" if ok2 else "BAD",
              json.dumps({k: round(v, 4) for k, v in row.items()}), flush=True)
    return bad


def autograd_cases():
    bad = 0
    torch.backends.cudnn.allow_tf32 = False
    for name, fn in (("pointwise_conv", fused.pointwise_conv), ("conv_bias", fused.conv_bias)):
        for (B, C, O, N) in [(4, 64, 256, 512), (2, 130, 48, 300), (4, 512, 1024, 256)]:
            x = torch.randn(B, C, N, device=dev, requires_grad=True); w = (torch.randn(O, C, 1, device=dev) / C ** 0.5).requires_grad_()
            bs = torch.randn(O, device=dev, requires_grad=True); g = torch.randn(B, O, N, device=dev)
            y = fn(x, w, bs); y.backward(g)
            got = [y.detach(), x.grad.clone(), w.grad.clone(), bs.grad.clone()]
            x.grad = w.grad = bs.grad = None
            y2 = F.conv1d(x, w, bs); y2.backward(g)
            ref = [y2.detach(), x.grad, w.grad, bs.grad]
            errs = [float((a - r).abs().max() / r.abs().max()) for a, r in zip(got, ref)]
            ok = all(e < 3e-3 for e in errs)
            bad += not ok
            print("autograd", name, (B, C, O, N), ["%.1e" % e for e in errs], "OK" if ok else "BAD", flush=True)
    torch.backends.cudnn.allow_tf32 = True
    return bad


if __name__ == "__main__":
    b1 = onehot()
    b2 = rand_cases()
    b3 = bias_pieces() + wgrad_cases()
    b4 = autograd_cases()
    print("BAD one-hot %d, random %d, bias %d, autograd %d" % (b1, b2, b3, b4))
    if "--time" in sys.argv and b1 + b2 == 0:
        timings()
