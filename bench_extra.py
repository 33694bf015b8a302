"""Secondary measurements printed inside bench.py's JSON line under "extra" (N=1 only): the other
operators of the hot path at the sizes the reference's models use them, each next to
the REFERENCE's own CUDA kernel recompiled for sm_100a (oracle/_ref/libref_ops.so — a checker/baseline,
never on the product path).  CUDA events, 3 warm-up + `iters` timed calls, median.
"""
import torch


def _time(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def vrcnet_census(dev, g, have_ref):
    """Every call into the operator library that one VRCNet training step makes (forward + backward), at the
    shapes of cfgs/vrcnet.yaml with B=32 (the ops see 2B=64 after vrcnet.py:452-454) — SURVEY.md §3.1 / §A3:
    5 FPS, 9 gathers, 3 groupings, 3 three_nn, 3 three_interpolate, 4 Chamfer forward+backward, and the backward
    scatters of the feature gathers / groupings / interpolations.  The dense layers between them (cuDNN / cuBLAS
    through PyTorch) are outside the operator boundary and not part of this number.  Ours through the C ABI with
    preallocated outputs; the reference's kernels through oracle/ref_cuda (which allocates its outputs the way the
    reference's Python does)."""
    import mm3d_pn2 as mm  # noqa: F401  (ensures the package is importable from the mirror)
    from mvp_benchmark_b200 import _lib
    from oracle import ref_cuda
    L, P = _lib.lib, _lib.ptr
    R = lambda *s: torch.rand(*s, device=dev, generator=g)  # noqa: E731
    N = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
    I = lambda hi, *s: torch.randint(0, hi, s, device=dev, generator=g, dtype=torch.int32)  # noqa: E731
    E = lambda *s: torch.empty(*s, device=dev)  # noqa: E731
    EI = lambda *s: torch.empty(*s, device=dev, dtype=torch.int32)  # noqa: E731
    S = _lib.stream_of(R(1))
    ours, ref, kinds, rkinds = [], [], [], []

    class _Tagged(list):  # list whose append also records the op class of the entry
        def __init__(self, tags):
            super().__init__()
            self.tags, self.tag = tags, None

        def append(self, fn):
            super().append(fn)
            self.tags.append(self.tag)

    ours, ref = _Tagged(kinds), _Tagged(rkinds)

    def tag(name):
        ours.tag = ref.tag = name

    def fps(b, n, m):
        tag("fps")
        x, o = R(b, n, 3), EI(b, m)
        ours.append(lambda: L.mvp_furthest_point_sampling(b, n, m, P(x), None, P(o), S))
        ref.append(lambda: ref_cuda.furthest_point_sample(x, m))

    def gather(b, c, n, m, grad):
        tag("gather_points fwd+bwd")
        f, i, o = N(b, c, n), I(n, b, m), E(b, c, m)
        ours.append(lambda: L.mvp_gather_points(b, c, n, m, P(f), P(i), P(o), S))
        ref.append(lambda: ref_cuda.gather_points(f, i))
        if grad:
            go, gp = N(b, c, m), E(b, c, n)
            ws = _lib.workspace(L.mvp_scatter_workspace_bytes(b, n, m), dev)
            ours.append(lambda: L.mvp_gather_points_grad_ws(b, c, n, m, P(go), P(i), P(gp), P(ws), ws.numel(), S))
            ref.append(lambda: ref_cuda.gather_points_grad(go, i, n))

    def group(b, c, n, p, s_):
        tag("grouping_operation fwd+bwd")
        f, i, o = N(b, c, n), I(n, b, p, s_), E(b, c, p, s_)
        go, gp = N(b, c, p, s_), E(b, c, n)
        ours.append(lambda: L.mvp_group_points(b, c, n, p, s_, P(f), P(i), P(o), S))
        ws = _lib.workspace(L.mvp_scatter_workspace_bytes(b, n, p * s_), dev)
        ours.append(lambda: L.mvp_group_points_grad_ws(b, c, n, p, s_, P(go), P(i), P(gp), P(ws), ws.numel(), S))
        ref.append(lambda: ref_cuda.group_points(f, i))
        ref.append(lambda: ref_cuda.group_points_grad(go, i, n))

    def unpool(b, c, m, n):  # three_nn (n targets from m sources) + three_interpolate forward / backward
        tag("three_nn + three_interpolate fwd+bwd")
        u, k, d, i = R(b, n, 3), R(b, m, 3), E(b, n, 3), EI(b, n, 3)
        f, w, o, go, gp = N(b, c, m), R(b, n, 3), E(b, c, n), N(b, c, n), E(b, c, m)
        i3 = I(m, b, n, 3)
        wn = _lib.workspace(L.mvp_three_nn_workspace_bytes(b, n, m), dev)
        ours.append(lambda: L.mvp_three_nn_ws(b, n, m, P(u), P(k), P(d), P(i), P(wn), wn.numel(), S))
        ours.append(lambda: L.mvp_three_interpolate(b, c, m, n, P(f), P(i3), P(w), P(o), S))
        ws = _lib.workspace(L.mvp_scatter_workspace_bytes(b, m, 3 * n), dev)
        ours.append(lambda: L.mvp_three_interpolate_grad_ws(b, c, n, m, P(go), P(i3), P(w), P(gp), P(ws), ws.numel(), S))
        ref.append(lambda: ref_cuda.three_nn(u, k))
        ref.append(lambda: ref_cuda.three_interpolate(f, i3, w))
        ref.append(lambda: ref_cuda.three_interpolate_grad(go, i3, w, m))

    def chamfer(b, n, m):
        tag("chamfer fwd+bwd")
        a, c_ = R(b, n, 3), R(b, m, 3)
        d1, d2, i1, i2 = E(b, n), E(b, m), EI(b, n), EI(b, m)
        g1, g2, gx = R(b, n), R(b, m), E(b * (n + m) * 3)
        ws = _lib.workspace(L.mvp_chamfer_forward_workspace_bytes(b, n, m), dev)
        ours.append(lambda: L.mvp_chamfer_forward(b, n, m, P(a), P(c_), P(d1), P(d2), P(i1), P(i2), P(ws), ws.numel(), S))
        ours.append(lambda: L.mvp_chamfer_backward(b, n, m, P(a), P(c_), P(g1), P(g2), P(i1), P(i2), P(gx[:b * n * 3]),
                                                   P(gx[b * n * 3:]), S))

        def ref_cd():
            o1, o2, j1, j2 = ref_cuda.chamfer_forward(a, c_)
            ref_cuda.chamfer_backward(a, c_, g1, g2, j1, j2)
        ref.append(ref_cd)

    fps(32, 2048, 2048), gather(32, 3, 2048, 2048, False)                         # vrcnet.py:451
    for n, c in ((3072, 64), (1536, 128), (768, 256)):                            # vrcnet.py:255-273, model_utils.py:88-110
        fps(64, n, n // 2), gather(64, 3, n, n // 2, False), gather(64, c, n, 10 * (n // 2), True), group(64, c, n, n // 2, 1)
    for c, m in ((512, 384), (256, 768), (128, 1536)):                            # vrcnet.py:288-292
        unpool(64, c, m, 2 * m)
    fps(64, 3072, 2048), gather(64, 3, 3072, 2048, False), gather(64, 64, 3072, 2048, True)   # vrcnet.py:380-383
    for m in (1024, 3072, 2048, 2048):                                            # vrcnet.py:509-512
        chamfer(64, 2048, m)

    def run_all(fns):
        for f in fns:
            rc = f()
            if isinstance(rc, int) and rc != 0:
                raise RuntimeError(f"operator call failed with code {rc}")

    res = {"calls": len(ours), "ours_ms": _time(lambda: run_all(ours), 5, 2)}
    if have_ref:
        res["ref_cuda_ms"] = _time(lambda: run_all(ref), 5, 2)
        res["speedup_vs_ref_cuda"] = res["ref_cuda_ms"] / res["ours_ms"]
    by = {}
    for fns, tags, col in ((ours, kinds, "ours_ms"),) + (((ref, rkinds, "ref_cuda_ms"),) if have_ref else ()):
        for name in dict.fromkeys(tags):
            sel = [f for f, t in zip(fns, tags) if t == name]
            by.setdefault(name, {})[col] = round(_time(lambda: run_all(sel), 5, 2), 4)
    res["by_operator"] = by
    # the same calls replayed from one CUDA graph: device time without per-call host work
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        S2 = _lib.stream_of(R(1))
        S.value = S2.value
        run_all(ours)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            run_all(ours)
    torch.cuda.synchronize()
    res["ours_cuda_graph_ms"] = _time(graph.replay, 5, 2)
    return res


def model_steps():
    """One training step of the reference's UNMODIFIED VRCNet (B = 32, cfgs/vrcnet.yaml) through tools/model_step.py:
    on the reference's kernels, on ours, and on ours with the opt-in model patches.  Needs the models staged under
    baseline/_ref/completion (oracle/build_ref.py, where /root/reference exists); returns None without them."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.abspath(__file__))
    if not os.path.isfile(os.path.join(root, "baseline", "_ref", "completion", "models", "vrcnet.py")):
        return None
    out = {"model": "vrcnet", "batch": 32, "step": "zero_grad + forward + backward + Adam (completion/train.py:122-142)"}
    for tag, extra in (("ref_cuda_ms", ["--ops", "ref"]), ("ours_ms", ["--ops", "ours"]),
                       ("ours_with_model_patches_ms", ["--ops", "ours", "--patch-knn"])):
        if tag == "ref_cuda_ms" and not os.path.isfile(os.path.join(root, "oracle", "_ref", "libref_ops.so")):
            continue
        try:
            p = subprocess.run([sys.executable, os.path.join(root, "tools", "model_step.py"), "--model", "vrcnet",
                                "--steps", "10", "--warmup", "3", *extra], capture_output=True, text=True, timeout=300)
            line = [l for l in p.stdout.splitlines() if l.startswith("MODEL_STEP ")]
            out[tag] = json.loads(line[-1][len("MODEL_STEP "):])["ms_per_step"] if line else None
        except Exception as e:  # secondary number
            out[tag] = None
            out["error"] = repr(e)
    if out.get("ref_cuda_ms") and out.get("ours_ms"):
        out["speedup_vs_ref_cuda"] = out["ref_cuda_ms"] / out["ours_ms"]
        if out.get("ours_with_model_patches_ms"):
            out["speedup_with_model_patches"] = out["ref_cuda_ms"] / out["ours_with_model_patches_ms"]
    return out


def run(dev, hbm_gbs=None):
    import json
    import os

    import mvp_benchmark_b200
    mvp_benchmark_b200.install()
    import metrics
    import mm3d_pn2 as mm
    from oracle import ref_cuda
    have_ref = ref_cuda.available()
    if hbm_gbs is None:
        try:
            hbm_gbs = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            hbm_gbs = 6650.0
    out = {"hbm_peak_gbs": hbm_gbs, "reference_cuda_available": have_ref}
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    R = lambda *s: torch.rand(*s, device=dev, generator=g)  # noqa: E731

    def entry(name, ours_ms, ref_ms, unit_count, unit, alg_bytes=None):
        e = {"ours_ms": ours_ms, "ours_per_s": unit_count / (ours_ms * 1e-3), "unit": unit}
        if ref_ms is not None:
            e["ref_cuda_ms"] = ref_ms
            e["speedup_vs_ref_cuda"] = ref_ms / ours_ms
        if alg_bytes is not None:
            e["algorithmic_GBps"] = alg_bytes / (ours_ms * 1e-3) / 1e9
            e["hbm_frac"] = e["algorithmic_GBps"] / hbm_gbs
        out[name] = e

    # ---- Chamfer at the headline size, reference CUDA kernels for comparison (fwd+bwd)
    b, n = 32, 16384
    x1, x2, g1, g2 = R(b, n, 3), R(b, n, 3), R(b, n), R(b, n)
    cd = metrics.cd()

    def ours_cd():
        a, c = x1.detach().requires_grad_(True), x2.detach().requires_grad_(True)
        o1, o2, _, _ = cd(a, c)
        torch.autograd.backward([o1, o2], [g1, g2])

    def ref_cd():
        o1, o2, j1, j2 = ref_cuda.chamfer_forward(x1, x2)
        ref_cuda.chamfer_backward(x1, x2, g1, g2, j1, j2)

    entry("chamfer_fwd_bwd_32x16384x16384", _time(ours_cd, 5), _time(ref_cd, 5) if have_ref else None,
          float(b) * n * n, "point-pairs/s", 52.0 * b * 2 * n)

    # ---- Chamfer at the VRCNet sizes (B=64: 2048 x {1024, 3072, 2048, 2048}, vrcnet.py:509-512)
    gt = R(64, 2048, 3)
    outs = [R(64, k, 3) for k in (1024, 3072, 2048, 2048)]
    gs = [(R(64, 2048), R(64, k)) for k in (1024, 3072, 2048, 2048)]

    def ours_cd4():
        for o, (ga, gb) in zip(outs, gs):
            a, c = gt.detach().requires_grad_(True), o.detach().requires_grad_(True)
            p, q, _, _ = cd(a, c)
            torch.autograd.backward([p, q], [ga, gb])

    def ref_cd4():
        for o, (ga, gb) in zip(outs, gs):
            p, q, j1, j2 = ref_cuda.chamfer_forward(gt, o)
            ref_cuda.chamfer_backward(gt, o, ga, gb, j1, j2)

    entry("chamfer_fwd_bwd_vrcnet_4calls_B64", _time(ours_cd4), _time(ref_cd4) if have_ref else None,
          64.0 * 2048 * (1024 + 3072 + 2048 + 2048), "point-pairs/s")

    # ---- BASELINE config C2 with the geometry PCN has at random initialisation: the ground truth fills the unit cube,
    # the prediction is a small blob inside it (completion/models/pcn.py:99-100: CD(gt 16384, coarse 1024) +
    # CD(gt 16384, fine 16384)) — nearly every ground-truth point is far outside the prediction's grid and is finished
    # by the bounding-box hierarchy of chamfer_rest.cu.  Forward only (the backward does not depend on the geometry).
    gt16 = R(32, 16384, 3)
    for tag, mm_ in (("coarse1024", 1024), ("fine16384", 16384)):
        blob = (0.5 + 0.02 * torch.randn(32, mm_, 3, device=dev, generator=g)).contiguous()
        entry(f"chamfer_fwd_pcn_init_32x16384_{tag}", _time(lambda: cd(gt16, blob), 10),
              _time(lambda: ref_cuda.chamfer_forward(gt16, blob), 5) if have_ref else None,
              32.0 * 16384 * mm_, "point-pairs/s", 20.0 * 32 * (16384 + mm_))
    # ... and two more hostile geometries at the headline size: disjoint clouds, a sphere surface
    far = (R(32, 16384, 3) + 3.0).contiguous()
    entry("chamfer_fwd_disjoint_32x16384x16384", _time(lambda: cd(x1, far), 10),
          _time(lambda: ref_cuda.chamfer_forward(x1, far), 5) if have_ref else None,
          32.0 * 16384 * 16384, "point-pairs/s", 20.0 * 32 * 2 * 16384)
    sp = [torch.nn.functional.normalize(torch.randn(32, 16384, 3, device=dev, generator=g), dim=2) * 0.5 + 0.5 for _ in range(2)]
    entry("chamfer_fwd_sphere_32x16384x16384", _time(lambda: cd(sp[0].contiguous(), sp[1].contiguous()), 10), None,
          32.0 * 16384 * 16384, "point-pairs/s", 20.0 * 32 * 2 * 16384)

    # ---- EMD, config C5: B=64, n=8192, eps 0.005, 50 rounds (forward; the backward is a trivial gather)
    e1, e2 = R(64, 8192, 3), R(64, 8192, 3)
    emd = metrics.emd()
    entry("emd_forward_64x8192_iters50", _time(lambda: emd(e1, e2, 0.005, 50), 3, 1),
          _time(lambda: ref_cuda.emd_forward(e1, e2, 0.005, 50), 3, 1) if have_ref else None,
          64.0 * 8192 * 8192, "point-pairs/s", 32.0 * 64 * 8192)
    e1l, e2l = e1[:16].contiguous(), e2[:16].contiguous()
    entry("emd_forward_16x8192_iters3000", _time(lambda: emd(e1l, e2l, 0.005, 3000), 2, 1),
          _time(lambda: ref_cuda.emd_forward(e1l, e2l, 0.005, 3000), 2, 1) if have_ref else None,
          16.0 * 8192 * 8192, "point-pairs/s", 32.0 * 16 * 8192)
    e1s, e2s = R(32, 2048, 3), R(32, 2048, 3)
    entry("emd_forward_32x2048_iters50", _time(lambda: emd(e1s, e2s, 0.005, 50), 5, 2),
          _time(lambda: ref_cuda.emd_forward(e1s, e2s, 0.005, 50), 5, 2) if have_ref else None,
          32.0 * 2048 * 2048, "point-pairs/s", 32.0 * 32 * 2048)

    # ---- mm3d knn (QueryAndGroup's neighbourhood when max_radius is None): centres = the cloud
    for (bb, nn, pp, kk) in [(64, 2048, 512, 16), (32, 4096, 1024, 32), (16, 8192, 2048, 64), (16, 4096, 1024, 100)]:
        x, c = R(bb, nn, 3), R(bb, pp, 3)
        entry(f"knn_{bb}x{nn}_centres{pp}_k{kk}", _time(lambda: mm.knn(kk, x, c, False), 5),
              _time(lambda: ref_cuda.knn(kk, x, c), 5) if have_ref else None, float(bb) * pp * nn, "point-pairs/s")

    # ---- FPS at the VRCNet sizes
    for (bb, nn, mm_) in [(32, 2048, 2048), (64, 3072, 2048), (64, 3072, 1536), (64, 1536, 768), (64, 768, 384)]:
        x = R(bb, nn, 3)
        entry(f"fps_{bb}x{nn}to{mm_}", _time(lambda: mm.furthest_point_sample(x, mm_), 5),
              _time(lambda: ref_cuda.furthest_point_sample(x, mm_), 5) if have_ref else None,
              float(bb) * mm_, "selected points/s", 12.0 * bb * nn + 4.0 * bb * mm_)

    # ---- bandwidth ops: timed through the C ABI with preallocated outputs (the Python layer adds ~30 us of host
    # work per call, which is not what a roofline fraction should measure); sizes are VRCNet's
    from mvp_benchmark_b200 import _lib
    L, P, S = _lib.lib, _lib.ptr, _lib.stream_of

    def gather_case(tag, bb, c, nn, mp):
        feat = torch.randn(bb, c, nn, device=dev, generator=g)
        idx = torch.randint(0, nn, (bb, mp), device=dev, generator=g, dtype=torch.int32)
        out = torch.empty(bb, c, mp, device=dev)
        go = torch.randn(bb, c, mp, device=dev, generator=g)
        gp = torch.empty(bb, c, nn, device=dev)
        fwd = lambda: _lib.check(L.mvp_gather_points(bb, c, nn, mp, P(feat), P(idx), P(out), S(feat)), "gather")  # noqa: E731
        ws = _lib.workspace(L.mvp_scatter_workspace_bytes(bb, nn, mp), dev)
        bwd = lambda: _lib.check(L.mvp_gather_points_grad_ws(bb, c, nn, mp, P(go), P(idx), P(gp), P(ws), ws.numel(), S(go)),  # noqa: E731
                                 "gather grad")
        entry(f"gather_points_{tag}", _time(fwd), _time(lambda: ref_cuda.gather_points(feat, idx)) if have_ref else None,
              float(bb) * c * mp, "elements/s", 4.0 * bb * mp * (2 * c + 1))
        entry(f"gather_points_grad_{tag}", _time(bwd),
              _time(lambda: ref_cuda.gather_points_grad(go, idx, nn)) if have_ref else None,
              float(bb) * c * mp, "elements/s", 4.0 * bb * mp * (2 * c + 1) + 4.0 * bb * c * nn)

    gather_case("64x64x3072_to_15360", 64, 64, 3072, 15360)
    gather_case("64x128x1536_to_7680", 64, 128, 1536, 7680)
    gather_case("64x256x768_to_3840", 64, 256, 768, 3840)

    def interp_case(tag, bb, c, m_, n_):
        f = torch.randn(bb, c, m_, device=dev, generator=g)
        i3 = torch.randint(0, m_, (bb, n_, 3), device=dev, generator=g, dtype=torch.int32)
        w = R(bb, n_, 3)
        out = torch.empty(bb, c, n_, device=dev)
        go = torch.randn(bb, c, n_, device=dev, generator=g)
        gp = torch.empty(bb, c, m_, device=dev)
        fwd = lambda: _lib.check(L.mvp_three_interpolate(bb, c, m_, n_, P(f), P(i3), P(w), P(out), S(f)), "interp")  # noqa: E731
        ws = _lib.workspace(L.mvp_scatter_workspace_bytes(bb, m_, 3 * n_), dev)
        bwd = lambda: _lib.check(L.mvp_three_interpolate_grad_ws(bb, c, n_, m_, P(go), P(i3), P(w), P(gp), P(ws), ws.numel(),  # noqa: E731
                                                                 S(go)), "interp grad")
        entry(f"three_interpolate_{tag}", _time(fwd),
              _time(lambda: ref_cuda.three_interpolate(f, i3, w)) if have_ref else None, float(bb) * c * n_, "elements/s",
              4.0 * bb * n_ * (2 * c + 6))
        entry(f"three_interpolate_grad_{tag}", _time(bwd),
              _time(lambda: ref_cuda.three_interpolate_grad(go, i3, w, m_)) if have_ref else None, float(bb) * c * n_, "elements/s",
              4.0 * bb * n_ * (2 * c + 6) + 4.0 * bb * c * m_)

    interp_case("64x128x1536_to_3072", 64, 128, 1536, 3072)
    interp_case("64x256x768_to_1536", 64, 256, 768, 1536)
    interp_case("64x512x384_to_768", 64, 512, 384, 768)
    u, k = R(64, 3072, 3), R(64, 1536, 3)
    entry("three_nn_64x3072_from_1536", _time(lambda: mm.three_nn(u, k)),
          _time(lambda: ref_cuda.three_nn(u, k)) if have_ref else None, 64.0 * 3072 * 1536, "point-pairs/s",
          12.0 * 64 * (3072 + 1536) + 24.0 * 64 * 3072)
    # ---- the kNN the MODELS run in torch (matmul expansion + topk, model_utils.py:242-259),
    # restated here, against the fused exact search (mvp_knn_points) at VRCNet's sizes (B = 2*32 after vrcnet.py:452)
    from mvp_benchmark_b200 import fused

    def torch_knn_point(pk, point_input, point_output):
        m_, n_ = point_output.size(1), point_input.size(1)
        inner = -2 * torch.matmul(point_output, point_input.transpose(2, 1).contiguous())
        xx = torch.sum(point_output ** 2, dim=2, keepdim=True).repeat(1, 1, n_)
        yy = torch.sum(point_input ** 2, dim=2, keepdim=False).unsqueeze(1).repeat(1, m_, 1)
        return (-xx - inner - yy).topk(k=pk, dim=-1)

    for tag, n_, m_, k_ in (("self_64x3072_k16", 3072, 3072, 16), ("self_64x1536_k16", 1536, 1536, 16),
                            ("point_64x1536_from_3072_k10", 1536, 3072, 10), ("point_64x384_from_768_k10", 384, 768, 10)):
        cl = R(64, m_, 3)
        qs = cl if tag.startswith("self") else R(64, n_, 3)
        e = {"ours_ms": _time(lambda: fused.knn_points(k_, cl, qs), 5, 2),
             "torch_matmul_topk_ms": _time(lambda: torch_knn_point(k_, cl, qs), 3, 1), "unit": "point-pairs/s"}
        e["ours_per_s"] = 64.0 * n_ * m_ / (e["ours_ms"] * 1e-3)
        e["speedup_vs_torch_formula"] = e["torch_matmul_topk_ms"] / e["ours_ms"]
        out["knn_points_" + tag] = e
    # ---- the fused caller-side operators of round 2 against the torch / operator sequences they replace
    def fused_entry(name, ours_fn, torch_fn, note):
        out[name] = {"ours_ms": _time(ours_fn, 5, 2), "torch_sequence_ms": _time(torch_fn, 3, 1), "replaces": note}
        out[name]["speedup"] = out[name]["torch_sequence_ms"] / out[name]["ours_ms"]

    ff = torch.randn(64, 64, 3072, device=dev, generator=g)
    fi = torch.randint(0, 3072, (64, 1536, 10), device=dev, generator=g, dtype=torch.int32)
    fused_entry("gather_max_64x64x3072_to_1536x10", lambda: fused.gather_max(ff, fi),
                lambda: torch.max(mm.gather_points(ff, fi.view(64, -1)).view(64, 64, 1536, 10), 3),
                "gather_points + view + torch.max (model_utils.py:97-102)")
    yy = torch.randn(64, 16, 3072, device=dev, generator=g)
    ww = torch.randn(64, 2, 20, 3072, device=dev, generator=g)
    wi = torch.randint(0, 3072, (64, 3072, 20), device=dev, generator=g, dtype=torch.int32)
    wit = wi.transpose(1, 2).contiguous()
    fused_entry("neighbor_weighted_sum_64x16x3072_k20", lambda: fused.neighbor_weighted_sum(yy, wi, ww),
                lambda: torch.sum(ww.repeat(1, 8, 1, 1) * mm.grouping_operation(yy, wit), dim=2),
                "grouping_operation + repeat + mul + sum (vrcnet.py:49-52)")
    sc = torch.randn(32, 2048, 2048, device=dev, generator=g)
    fused_entry("topk_rows_32x2048x2048_k16", lambda: fused.topk_rows(sc, 16), lambda: sc.topk(16, dim=-1),
                "torch.topk on the feature-space score matrix (model_utils.py:246)")
    fx = torch.randn(32, 24, 3072, device=dev, generator=g)

    def knn_original():
        inner = -2 * torch.matmul(fx.transpose(2, 1).contiguous(), fx)
        xx = torch.sum(fx ** 2, dim=1, keepdim=True)
        return (-xx - inner - xx.transpose(2, 1).contiguous()).topk(k=16, dim=-1)[1]
    fused_entry("feature_knn_32x24x3072_k16", lambda: fused.topk_rows_sqdist(torch.matmul(fx.transpose(2, 1).contiguous(), fx), torch.sum(fx ** 2, dim=1), 16),
                knn_original, "knn on features: matmul + three elementwise passes + torch.topk (model_utils.py:242-247)")
    fp = R(64, 3072, 3)
    fused_entry("fps_gather_64x3072_to_1536", lambda: fused.fps_gather(fp, 1536),
                lambda: mm.gather_points(fp.transpose(1, 2).contiguous(), mm.furthest_point_sample(fp, 1536)).transpose(1, 2).contiguous(),
                "furthest_point_sample + transpose + gather_points + transpose (model_utils.py:91-93)")
    # ---- the 1x1 layers (csrc/pointwise.cu): tcgen05 contraction with the bias in its epilogue, against cuDNN's TF32
    # convolution + torch's bias add; the bias kernels of the wide layers against torch's add / sum
    import torch.nn.functional as F
    for (pb, pc, po, pn) in ((64, 64, 256, 3072), (64, 128, 256, 2048)):
        px = torch.randn(pb, pc, pn, device=dev, generator=g)
        pw = torch.randn(po, pc, device=dev, generator=g) / pc ** 0.5
        pbias = torch.randn(po, device=dev, generator=g)
        pg = torch.randn(pb, po, pn, device=dev, generator=g)
        pwt, pw3 = pw.t().contiguous(), pw.view(po, pc, 1)
        tag = "%dx%dto%dx%d" % (pb, pc, po, pn)
        fused_entry("pointwise_conv_forward_" + tag, lambda: fused._pointwise_conv_raw(px, pw, pbias),
                    lambda: F.conv1d(px, pw3, pbias), "nn.Conv1d / nn.Conv2d(kernel 1) forward: cuDNN TF32 + bias add")
        out["pointwise_conv_forward_" + tag]["algorithmic_GBps"] = 4.0 * pb * (pc + po) * pn / out["pointwise_conv_forward_" + tag]["ours_ms"] / 1e6
        fused_entry("pointwise_conv_input_grad_" + tag, lambda: fused._pointwise_conv_raw(pg, pwt, None),
                    lambda: torch.ops.aten.convolution_backward(pg, px, pw3, None, [1], [0], [1], False, [0], 1, [True, False, False]),
                    "the layer's input gradient: cuDNN dgrad")
        fused_entry("pointwise_wgrad_bias_" + tag, lambda: fused._pointwise_wgrad_raw(pg, px, True),
                    lambda: (torch.ops.aten.convolution_backward(pg, px, pw3, None, [1], [0], [1], False, [0], 1, [False, True, False]), pg.sum((0, 2))),
                    "the layer's weight + bias gradients: cuDNN wgrad + torch.sum")
    by = torch.randn(64, 1024, 2048, device=dev, generator=g)
    bb = torch.randn(1024, device=dev, generator=g)
    fused_entry("bias_add_64x1024x2048", lambda: fused.bias_add_(by, bb), lambda: by.add_(bb.view(1, -1, 1)),
                "the bias add of a wide 1x1 layer (torch's broadcasting add_)")
    out["bias_add_64x1024x2048"]["algorithmic_GBps"] = 2 * 4.0 * by.numel() / out["bias_add_64x1024x2048"]["ours_ms"] / 1e6
    fused_entry("channel_sum_64x1024x2048", lambda: fused.channel_sum(by), lambda: by.sum((0, 2)),
                "the bias gradient of a wide 1x1 layer (torch.sum over clouds and points)")
    out["channel_sum_64x1024x2048"]["algorithmic_GBps"] = 4.0 * by.numel() / out["channel_sum_64x1024x2048"]["ours_ms"] / 1e6
    del by
    mx = torch.randn(32, 96, 3072, 16, device=dev, generator=g)
    fused_entry("max_last_32x96x3072x16", lambda: fused.max_last(mx), lambda: torch.max(mx, 3),
                "torch.max over the k neighbours (ecg.py:64)")
    out["max_last_32x96x3072x16"]["algorithmic_GBps"] = 4.0 * mx.numel() / out["max_last_32x96x3072x16"]["ours_ms"] / 1e6
    del mx
    out["vrcnet_step_operator_census"] = vrcnet_census(dev, g, have_ref)
    steps = model_steps()
    if steps is not None:
        out["vrcnet_training_step"] = steps
    xyz, ctr = R(32, 2048, 3), R(32, 102, 3)
    entry("ball_query_32x2048_102centres_ns12", _time(lambda: mm.ball_query(0, 0.0774596669, 12, xyz, ctr)),
          _time(lambda: ref_cuda.ball_query(0, 0.0774596669, 12, xyz, ctr)) if have_ref else None, 32.0 * 102,
          "centres/s", 12.0 * 32 * (2048 + 102) + 4.0 * 32 * 102 * 12)
    return out
