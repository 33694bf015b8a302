"""GPU: the reference-facing Python layer (autograd Functions / Modules named as in the reference)
drives the same kernels — shapes, dtypes, non-differentiable outputs, backward arities, streams."""
import numpy as np
import pytest
import torch

import _cases
import _data
from _impls import OracleImpl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cpu():
    return OracleImpl()


def T(a, dev, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t.requires_grad_(True) if grad else t


def test_cd_module_forward_backward(ops, cuda, cpu):
    metrics, _ = ops
    x1, x2 = _data.uniform(4, 100, 0), _data.uniform(4, 200, 1)      # shapes of unit_test.py:15-16
    a, c = T(x1, cuda, True), T(x2, cuda, True)
    dist1, dist2, idx1, idx2 = metrics.cd()(a, c)
    assert dist1.shape == (4, 100) and dist2.shape == (4, 200) and idx1.dtype == torch.int32 and idx2.dtype == torch.int32
    assert not idx1.requires_grad and dist1.requires_grad
    w1, w2 = torch.rand_like(dist1), torch.rand_like(dist2)
    ((dist1 * w1).sum() + (dist2 * w2).sum()).backward()
    od1, od2, oi1, oi2 = cpu.chamfer_forward(x1, x2)
    _cases.eq(dist1.detach().cpu().numpy(), od1, "dist1"), _cases.eq(idx2.cpu().numpy(), oi2, "idx2")
    g1, g2 = cpu.chamfer_backward(x1, x2, w1.cpu().numpy(), w2.cpu().numpy(), oi1, oi2)
    _cases.close(a.grad.cpu().numpy(), g1, "gradxyz1"), _cases.close(c.grad.cpu().numpy(), g2, "gradxyz2")
    # the reference's own assertion against its pure-torch Chamfer (unit_test.py:22-33)
    from metrics.CD import chamfer_python
    m1, m2, mi1, mi2 = chamfer_python.distChamfer(a.detach(), c.detach())
    assert (torch.mean((dist1 - m1) ** 2) + torch.mean((dist2 - m2) ** 2)).item() < 1e-8
    assert (idx1 == mi1).all() and (idx2 == mi2).all()
    f, p1, p2 = metrics.fscore(dist1, dist2)
    assert f.shape == (4,)


def test_cd_noncontiguous_and_only_one_output_used(ops, cuda):
    metrics, _ = ops
    a = torch.rand(2, 3, 300, device=cuda).transpose(1, 2).requires_grad_(True)     # (2,300,3) non-contiguous
    c = torch.rand(2, 150, 3, device=cuda)
    d1, _, _, _ = metrics.cd()(a, c)
    d1.mean().backward()
    assert a.grad is not None and a.grad.shape == (2, 300, 3) and torch.isfinite(a.grad).all()


def test_calc_cd_formula_as_the_models_use_it(ops, cuda):
    """completion/model_utils.py:67-77 on top of the op."""
    metrics, _ = ops
    out, gt = torch.rand(3, 512, 3, device=cuda), torch.rand(3, 700, 3, device=cuda)
    dist1, dist2, _, _ = metrics.cd()(gt, out)
    cd_p = (torch.sqrt(dist1).mean(1) + torch.sqrt(dist2).mean(1)) / 2
    cd_t = dist1.mean(1) + dist2.mean(1)
    ref = torch.cdist(gt.double(), out.double()) ** 2
    assert torch.allclose(cd_t.double(), ref.min(2)[0].mean(1) + ref.min(1)[0].mean(1), rtol=1e-5)
    assert cd_p.shape == (3,)


def test_emd_module(ops, cuda, cpu):
    metrics, _ = ops
    x1, x2 = _data.uniform(2, 1024, 3), _data.uniform(2, 1024, 4)
    a = T(x1, cuda, True)
    dist, asg = metrics.emd()(a, T(x2, cuda), 0.005, 50)
    assert dist.shape == (2, 1024) and asg.dtype == torch.int32 and not asg.requires_grad
    torch.sqrt(dist).mean(1).sum().backward()                       # calc_emd, model_utils.py:80-85
    od, oa = cpu.emd_forward(x1, x2, 0.005, 50)
    _cases.eq(asg.cpu().numpy(), oa, "assignment"), _cases.eq(dist.detach().cpu().numpy(), od, "dist")
    assert torch.isfinite(a.grad).all() and a.grad.abs().sum() > 0
    with pytest.raises(AssertionError):
        metrics.emd()(a, torch.rand(2, 2048, 3, device=cuda), 0.005, 5)     # emd_module.py:47
    with pytest.raises(RuntimeError, match="multiple of 1024"):
        metrics.emd()(torch.rand(1, 1000, 3, device=cuda), torch.rand(1, 1000, 3, device=cuda), 0.005, 5)


def test_sampling_chain_as_in_vrcnet(ops, cuda, cpu):
    """FPS -> gather -> group -> three_nn -> three_interpolate, as completion/model_utils.py:88-110,286-293."""
    _, mm = ops
    B, N, C, S = 4, 1536, 64, 768
    xyz = T(_data.uniform(B, N, 5), cuda)
    feat = torch.randn(B, C, N, device=cuda, requires_grad=True)
    p_idx = mm.furthest_point_sample(xyz, S)
    assert p_idx.dtype == torch.int32 and p_idx.shape == (B, S) and not p_idx.requires_grad
    _cases.eq(p_idx.cpu().numpy(), cpu.fps(xyz.cpu().numpy(), S), "fps")
    pts = mm.gather_points(xyz.transpose(1, 2).contiguous(), p_idx).transpose(1, 2).contiguous()
    assert torch.equal(pts, torch.gather(xyz, 1, p_idx.long()[..., None].expand(-1, -1, 3)))
    center = mm.grouping_operation(feat, p_idx.unsqueeze(2).contiguous()).view(B, -1, S)
    assert torch.equal(center, torch.gather(feat, 2, p_idx.long()[:, None, :].expand(-1, C, -1)))
    dist, idx = mm.three_nn(xyz, pts)
    assert dist.shape == (B, N, 3) and idx.dtype == torch.int32
    d2, i2 = cpu.three_nn(xyz.cpu().numpy(), pts.cpu().numpy())
    _cases.eq(idx.cpu().numpy(), i2, "three_nn idx")
    np.testing.assert_array_max_ulp(dist.cpu().numpy(), np.sqrt(d2), maxulp=1)
    dist = torch.max(dist, torch.ones(1, device=cuda) * 1e-10)
    w = (1.0 / dist) / torch.sum(1.0 / dist, 2, keepdim=True)
    up = mm.three_interpolate(center.contiguous(), idx, w.contiguous())
    assert up.shape == (B, C, N)
    up.square().sum().backward()
    assert feat.grad is not None and torch.isfinite(feat.grad).all() and feat.grad.abs().sum() > 0


def test_ball_query_group_knn_modules(ops, cuda, cpu):
    _, mm = ops
    xyz = T(_data.uniform(2, 1024, 6), cuda)
    new_xyz = xyz[:, :51].contiguous()
    idx = mm.ball_query(0, 0.0774596669, 6, xyz, new_xyz)           # model_utils.py:211
    _cases.eq(idx.cpu().numpy(), cpu.ball_query(0, 0.0774596669, 6, xyz.cpu().numpy(), new_xyz.cpu().numpy()), "bq")
    grouped = mm.grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
    assert grouped.shape == (2, 3, 51, 6)
    k_idx = mm.knn(8, xyz, new_xyz, False)
    assert k_idx.shape == (2, 8, 51)
    oi, _ = cpu.knn(8, xyz.cpu().numpy(), new_xyz.cpu().numpy())
    _cases.eq(k_idx.cpu().numpy(), oi.transpose(0, 2, 1), "knn")
    assert torch.equal(mm.knn(8, xyz.transpose(1, 2).contiguous(), new_xyz.transpose(1, 2).contiguous(), True), k_idx)
    feats = torch.randn(2, 5, 1024, device=cuda)
    out = mm.QueryAndGroup(0.1, 16)(xyz, new_xyz, feats)
    assert out.shape == (2, 8, 51, 16)
    out = mm.QueryAndGroup(None, 4, return_grouped_xyz=True)(xyz, new_xyz, feats)
    assert out[0].shape == (2, 8, 51, 4) and out[1].shape == (2, 3, 51, 4)
    s = mm.Points_Sampler([16, 16], ['D-FPS', 'F-FPS'], [512, -1])(xyz, feats)
    assert s.shape == (2, 32) and s[:, :16].max() < 512 and s[:, 16:].min() >= 512
    d = mm.furthest_point_sample_with_dist(torch.cdist(xyz[:, :200], xyz[:, :200]).contiguous(), 20)
    assert d.shape == (2, 20) and (d[:, 0] == 0).all()


def test_ops_follow_the_current_stream(ops, cuda):
    metrics, _ = ops
    s = torch.cuda.Stream(device=cuda)
    a, c = torch.rand(8, 2048, 3, device=cuda), torch.rand(8, 2048, 3, device=cuda)
    torch.cuda.synchronize()
    want = metrics.cd()(a, c)[0].clone()
    with torch.cuda.stream(s):
        got = metrics.cd()(a, c)[0]
    s.synchronize()
    assert torch.equal(got, want)


def test_ops_are_cuda_graph_capturable(ops, cuda):
    """No op synchronises the host or allocates behind torch's back: a forward+backward through Chamfer (grid
    path and the small-cloud path), EMD, FPS, gather and three_interpolate is captured once and replayed on new
    inputs; the replay must reproduce the eager results bit for bit (gradients: same accumulation class, 1e-5)."""
    metrics, mm = ops
    g = torch.Generator(device=cuda)
    g.manual_seed(3)
    R = lambda *s: torch.rand(*s, device=cuda, generator=g)  # noqa: E731
    x1, x2, small = R(2, 2048, 3), R(2, 1536, 3), R(2, 100, 3)
    e1, e2 = R(2, 1024, 3), R(2, 1024, 3)
    feat = R(2, 16, 2048)
    i3 = torch.randint(0, 2048, (2, 1536, 3), device=cuda, generator=g, dtype=torch.int32)
    w3 = R(2, 1536, 3)
    cd, emd = metrics.cd(), metrics.emd()

    def step():
        a, f = x1.detach().requires_grad_(True), feat.detach().requires_grad_(True)
        d1, d2, j1, j2 = cd(a, x2)
        s1, s2, _, _ = cd(small, a)
        ed, ea = emd(e1, e2, 0.005, 20)
        fi = mm.furthest_point_sample(a.detach(), 512)
        gathered = mm.gather_points(f, fi)
        interp = mm.three_interpolate(f, i3, w3)
        (d1.sum() + d2.sum() + s2.sum() + gathered.sum() + (interp * w3[..., 0].unsqueeze(1)).sum()).backward()
        return [d1.detach(), d2.detach(), j1, j2, s1.detach(), ed, ea, fi, gathered.detach(), interp.detach()], [a.grad, f.grad]

    side = torch.cuda.Stream(cuda)
    side.wait_stream(torch.cuda.current_stream(cuda))
    with torch.cuda.stream(side):
        step()  # warm-up outside capture (lazy kernel-attribute setup, allocator pools)
    torch.cuda.current_stream(cuda).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs, grads = step()
    for t in (x1, x2, small, e1, e2, feat, w3):  # new inputs, same buffers
        t.copy_(torch.rand(t.shape, device=cuda, generator=g))
    graph.replay()
    torch.cuda.synchronize(cuda)
    got = [t.clone() for t in outs], [t.clone() for t in grads]
    want = step()
    torch.cuda.synchronize(cuda)
    for a, b in zip(got[0], want[0]):
        assert torch.equal(a, b)
    for a, b in zip(got[1], want[1]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)


def _torch_knn_point(pk, point_input, point_output):
    """The formula of completion/model_utils.py:250-259, restated for the test (matmul expansion + topk)."""
    inner = -2 * torch.matmul(point_output, point_input.transpose(2, 1))
    xx = torch.sum(point_output ** 2, dim=2, keepdim=True)
    yy = torch.sum(point_input ** 2, dim=2).unsqueeze(1)
    return (-xx - inner - yy).topk(k=pk, dim=-1)


def test_model_patches_knn_and_knn_point(cuda, cpu):
    """Opt-in kNN replacements: same shapes / dtypes / sign conventions as
    model_utils.knn / knn_point, same neighbour sets as the matmul + topk formula away from near-ties,
    gradients through the returned distances, non-3-D inputs fall through to the original function."""
    import types
    from mvp_benchmark_b200 import model_patches as mp
    calls = []

    def orig_knn(x, k):
        calls.append("knn")
        return torch.zeros(x.size(0), x.size(2), k, dtype=torch.long, device=x.device)

    fake = types.SimpleNamespace(knn=orig_knn, knn_point=_torch_knn_point, knn_point_all=_torch_knn_point)
    assert mp.apply(fake) == 3 and mp.apply(fake) == 0   # (this stand-in module has no get_edge_features)
    B, N, M = 4, 1536, 768
    cloud = T(_data.uniform(B, N, 11), cuda)
    sub = cloud[:, :M].contiguous().requires_grad_(True)
    idx = fake.knn(cloud.transpose(1, 2).contiguous(), 16)
    assert idx.shape == (B, N, 16) and idx.dtype == torch.int64 and not calls
    _cases.eq(idx.cpu().numpy().astype(np.int32), cpu.o.knn_points(16, cloud.cpu().numpy())[1], "patched knn")
    negd, pidx = fake.knn_point(10, cloud, sub)
    assert negd.shape == (B, M, 10) and pidx.dtype == torch.int64 and negd.requires_grad
    ref_d, ref_i = _torch_knn_point(10, cloud, sub.detach())
    torch.testing.assert_close(negd.detach(), ref_d, atol=5e-6, rtol=0)
    agree = (pidx.sort(dim=2)[0] == ref_i.sort(dim=2)[0]).all(dim=2).float().mean().item()
    assert agree > 0.995, agree
    negd.sum().backward()
    assert sub.grad is not None and torch.isfinite(sub.grad).all()
    feat = torch.randn(2, 64, 100, device=cuda)            # feature-space kNN: the original's matrix, our row-wise top-k
    got = fake.knn(feat, 5)
    inner = -2 * torch.matmul(feat.transpose(2, 1).contiguous(), feat)
    xx = torch.sum(feat ** 2, dim=1, keepdim=True)
    want = (-xx - inner - xx.transpose(2, 1).contiguous()).topk(k=5, dim=-1)[1]
    assert not calls and got.dtype == torch.int64 and torch.equal(got, want)
    fake.knn(feat, 40)                                      # k > 32: the original
    assert calls == ["knn"]


def test_model_patches_get_edge_features(cuda):
    """The neighbour-feature gather of SA_module (model_utils.py:113-124) as one grouping_operation: same values bit
    for bit, contiguous (B, C, k, N), gradients equal to the advanced-indexing original's up to summation order."""
    import types
    from mvp_benchmark_b200 import model_patches as mp

    def original(x, idx):                                  # model_utils.py:113-124, restated
        batch_size, num_points, k = idx.size()
        idx = (idx + torch.arange(0, batch_size, device=x.device).view(-1, 1, 1) * num_points).view(-1)
        x = x.squeeze(2)
        num_dims = x.size(1)
        x = x.transpose(2, 1).contiguous()
        feature = x.view(batch_size * num_points, -1)[idx, :]
        return feature.view(batch_size, num_points, k, num_dims).permute(0, 3, 2, 1)

    fake = types.SimpleNamespace(get_edge_features=original)
    assert mp.apply(fake) == 1
    B, C, N, k = 3, 37, 1536, 16
    x0 = torch.randn(B, C, 1, N, device=cuda)
    idx = torch.randint(0, N, (B, N, k), device=cuda)
    a, b = x0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
    got, want = fake.get_edge_features(a, idx), original(b, idx)
    assert got.shape == (B, C, k, N) and got.is_contiguous() and torch.equal(got, want)
    g = torch.randn_like(got)
    got.backward(g), want.backward(g)
    torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("minus_center", [True, False])
def test_model_patches_get_graph_feature(cuda, minus_center):
    """ECG's / EF_expansion's neighbour-feature builder (model_utils.py:156-178) through one grouping_operation: the
    same (B, 2C, N, k) values bit for bit (given the same neighbours), contiguous, gradients equal up to summation order."""
    import types
    from mvp_benchmark_b200 import model_patches as mp

    def knn(x, k):                                         # model_utils.py:242-247, restated
        inner = -2 * torch.matmul(x.transpose(2, 1).contiguous(), x)
        xx = torch.sum(x ** 2, dim=1, keepdim=True)
        return (-xx - inner - xx.transpose(2, 1).contiguous()).topk(k=k, dim=-1)[1]

    def original(x, k=20, minus_center=True):              # model_utils.py:156-178, restated
        idx = knn(x, k=k)
        batch_size, num_points, _ = idx.size()
        idx = (idx + torch.arange(0, batch_size, device=x.device).view(-1, 1, 1) * num_points).view(-1)
        _, num_dims, _ = x.size()
        x = x.transpose(2, 1).contiguous()
        feature = x.view(batch_size * num_points, -1)[idx, :].view(batch_size, num_points, k, num_dims)
        x = x.view(batch_size, num_points, 1, num_dims).repeat(1, 1, k, 1)
        return torch.cat((x, feature - x if minus_center else feature), dim=3).permute(0, 3, 1, 2)

    fake = types.SimpleNamespace(get_graph_feature=original, knn=knn)
    assert mp.apply(fake) == 2
    B, C, N, k = 3, 24, 700, 16
    x0 = torch.randn(B, C, N, device=cuda)
    a, b = x0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
    got, want = fake.get_graph_feature(a, k, minus_center), original(b, k, minus_center)
    assert got.shape == (B, 2 * C, N, k) and got.is_contiguous() and torch.equal(got, want)
    g = torch.randn_like(got)
    got.backward(g), want.backward(g)
    torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("b,n,m", [(64, 2048, 1024), (3, 16384, 16384), (5, 7, 3), (2, 1000, 1)])
def test_chamfer_loss_epilogue(cuda, cpu, b, n, m):
    """fused.chamfer_loss against the oracle and against the torch formula of
    completion/model_utils.py:71-72, forward and backward, 1e-5 relative."""
    from mvp_benchmark_b200 import fused
    rng = np.random.default_rng(5)
    d1 = (rng.random((b, n), dtype=np.float32) ** 2 * 0.01).astype(np.float32)
    d2 = (rng.random((b, m), dtype=np.float32) ** 2 * 0.01).astype(np.float32)
    a, c = T(d1, cuda, grad=True), T(d2, cuda, grad=True)
    cd_p, cd_t = fused.chamfer_loss(a, c)
    want_p, want_t = cpu.o.chamfer_loss(d1, d2)
    np.testing.assert_allclose(cd_p.detach().cpu().numpy(), want_p, rtol=1e-5)
    np.testing.assert_allclose(cd_t.detach().cpu().numpy(), want_t, rtol=1e-5)
    wp, wt = torch.rand(b, device=cuda), torch.rand(b, device=cuda)
    ((cd_p * wp).sum() + (cd_t * wt).sum()).backward()
    a2, c2 = T(d1, cuda, grad=True), T(d2, cuda, grad=True)
    ref_p = (torch.sqrt(a2).mean(1) + torch.sqrt(c2).mean(1)) / 2
    ref_t = a2.mean(1) + c2.mean(1)
    ((ref_p * wp).sum() + (ref_t * wt).sum()).backward()
    torch.testing.assert_close(cd_p.detach(), ref_p.detach(), rtol=1e-5, atol=0)
    torch.testing.assert_close(a.grad, a2.grad, rtol=1e-5, atol=0)
    torch.testing.assert_close(c.grad, c2.grad, rtol=1e-5, atol=0)


def test_model_patches_calc_cd(ops, cuda):
    """calc_cd rebinding: Chamfer operator + fused epilogue, same returns as model_utils.calc_cd (with and without
    the F-score), gradients reach the prediction."""
    import types
    from mvp_benchmark_b200 import model_patches as mp
    metrics, _ = ops

    def original(output, gt, calc_f1=False):                # model_utils.py:67-77, restated
        dist1, dist2, _, _ = metrics.cd()(gt, output)
        cd_p = (torch.sqrt(dist1).mean(1) + torch.sqrt(dist2).mean(1)) / 2
        cd_t = dist1.mean(1) + dist2.mean(1)
        if calc_f1:
            return cd_p, cd_t, metrics.fscore(dist1, dist2)[0]
        return cd_p, cd_t

    fake = types.SimpleNamespace(calc_cd=original)
    assert mp.apply(fake) == 1
    gt = T(_data.uniform(4, 2048, 21), cuda)
    p1, p2 = T(_data.uniform(4, 1024, 22), cuda, grad=True), T(_data.uniform(4, 1024, 22), cuda, grad=True)
    got, want = fake.calc_cd(p1, gt, calc_f1=True), original(p2, gt, calc_f1=True)
    for g_, w_ in zip(got, want):
        torch.testing.assert_close(g_, w_, rtol=1e-5, atol=0)
    got[0].sum().backward(), want[0].sum().backward()
    torch.testing.assert_close(p1.grad, p2.grad, rtol=1e-4, atol=1e-7)


def test_three_nn_weights_chain(ops, cuda):
    """fused.three_nn_weights / the three_nn_upsampling patch against three_nn + the torch glue
    of completion/model_utils.py:286-293: same indices, weights within 2 ulp (bit-equal with the torch build the order of the sum was measured on)."""
    import types
    from mvp_benchmark_b200 import model_patches as mp
    _, mm = ops

    def original(target_points, source_points):            # model_utils.py:286-293, restated
        dist, idx = mm.three_nn(target_points, source_points)
        dist = torch.max(dist, torch.ones(1, device=dist.device) * 1e-10)
        norm = torch.sum((1.0 / dist), 2, keepdim=True)
        norm = norm.repeat(1, 1, 3)
        return idx, (1.0 / dist) / norm

    fake = types.SimpleNamespace(three_nn_upsampling=original)
    assert mp.apply(fake) == 1
    for kind, n, m in (("uniform", 3072, 1536), ("duplicates", 768, 384), ("lattice", 300, 100)):
        t, s = T(_data.cloud(kind, 4, n, 31), cuda), T(_data.cloud(kind, 4, m, 32), cuda)
        if kind != "uniform":
            s = t[:, :m].contiguous()                       # exact coincidences: distance 0 -> the 1e-10 clamp
        (gi, gw), (wi, ww) = fake.three_nn_upsampling(t, s), original(t, s)
        assert torch.equal(gi, wi) and gw.shape == (4, n, 3)
        # bit-equal with torch 2.11 (same IEEE operations, its order of adding the three reciprocals); 2 ulp otherwise
        torch.testing.assert_close(gw, ww, rtol=3e-7, atol=0)


def test_fps_gather_and_ball_query_group_chains(ops, cuda):
    """The remaining fused chains of SURVEY.md §8f row 2 against the sequences of operator calls they replace:
    FPS -> transpose -> gather_points -> transpose (completion/model_utils.py:91-93, :209-210; vrcnet.py:451 keeps the
    (B, 3, m) layout) and ball_query -> grouping_operation -> permute (:211-214).  Same bits forward; gradients equal to
    the operators' own scatters."""
    import types
    from mvp_benchmark_b200 import fused, model_patches as mp
    _, mm = ops
    for kind, b, n, m in (("uniform", 4, 2048, 512), ("lattice", 2, 700, 700), ("duplicates", 3, 5000, 300), ("uniform", 2, 20000, 64),
                          ("uniform", 70, 96, 33)):
        pts = T(_data.cloud(kind, b, n, 41), cuda)
        a, c = pts.clone().requires_grad_(True), pts.clone().requires_grad_(True)
        idx, out = fused.fps_gather(a, m)
        widx = mm.furthest_point_sample(c, m)
        want = mm.gather_points(c.transpose(1, 2).contiguous(), widx).transpose(1, 2).contiguous()
        assert torch.equal(idx, widx) and idx.dtype == torch.int32 and out.shape == (b, m, 3) and torch.equal(out, want)
        _, out_cf = fused.fps_gather(pts, m, channels_first=True)
        assert out_cf.shape == (b, 3, m) and torch.equal(out_cf, want.transpose(1, 2))
        g = torch.randn_like(out)
        out.backward(g), want.backward(g)
        torch.testing.assert_close(a.grad, c.grad, rtol=1e-5, atol=1e-6)
    for kind, b, n, p, r, ns in (("uniform", 4, 2048, 102, 0.0632455532, 8), ("lattice", 2, 1500, 33, 0.13, 24),
                                 ("uniform", 2, 3000, 257, 1e-6, 4), ("uniform", 3, 40, 9, 5.0, 50)):
        pcd = T(_data.cloud(kind, b, n, 42), cuda)
        centres = pcd[:, :p].contiguous()
        a, c = pcd.clone().requires_grad_(True), pcd.clone().requires_grad_(True)
        idx, grouped = fused.ball_query_group(0, r, ns, a, centres)
        widx = mm.ball_query(0, r, ns, c, centres)
        want = mm.grouping_operation(c.transpose(1, 2).contiguous(), widx).permute(0, 2, 3, 1).contiguous()
        assert torch.equal(idx, widx) and grouped.shape == (b, p, ns, 3) and torch.equal(grouped, want)
        g = torch.randn_like(grouped)
        grouped.backward(g), want.backward(g)
        torch.testing.assert_close(a.grad, c.grad, rtol=1e-5, atol=1e-5)

    # the two callers, patched, against their originals restated on the reference-facing operators
    def orig_knn_point(pk, point_input, point_output):
        return _torch_knn_point(pk, point_input, point_output)

    def orig_eps(feature_input, point_input, num_samples, k=10):   # model_utils.py:86-108, restated
        batch_size, feature_size, num_points = feature_input.size()
        p_idx = mm.furthest_point_sample(point_input, num_samples)
        point_output = mm.gather_points(point_input.transpose(1, 2).contiguous(), p_idx).transpose(1, 2).contiguous()
        pk = int(min(k, num_points))
        _, pn_idx = fake.knn_point(pk, point_input, point_output)
        pn_idx = pn_idx.detach().int()
        neighbor_feature = mm.gather_points(feature_input, pn_idx.view(batch_size, num_samples * pk)).view(
            batch_size, feature_size, num_samples, pk)
        neighbor_feature, _ = torch.max(neighbor_feature, 3)
        center_feature = mm.grouping_operation(feature_input, p_idx.unsqueeze(2)).view(batch_size, -1, num_samples)
        return torch.cat((center_feature, neighbor_feature), 1), p_idx, pn_idx, point_output

    # gather + max over the neighbours (fused.gather_max) against gather_points + torch.max, ties included
    for B_, C_, N_, M_, K_ in ((4, 64, 3072, 1536, 10), (2, 37, 700, 333, 16), (3, 5, 16384, 100, 3), (70, 3, 64, 64, 1)):
        feat = torch.randn(B_, C_, N_, device=cuda)
        feat = torch.round(feat * 4) / 4                    # exact ties: the first neighbour among equal maxima wins
        nidx = torch.randint(0, N_, (B_, M_, K_), device=cuda, dtype=torch.int32)
        a, c = feat.clone().requires_grad_(True), feat.clone().requires_grad_(True)
        got = fused.gather_max(a, nidx)
        want, _ = torch.max(mm.gather_points(c, nidx.view(B_, M_ * K_)).view(B_, C_, M_, K_), 3)
        assert torch.equal(got, want)
        g = torch.randn_like(got)
        got.backward(g), want.backward(g)
        torch.testing.assert_close(a.grad, c.grad, rtol=1e-5, atol=1e-5)

    fake = types.SimpleNamespace(edge_preserve_sampling=orig_eps, knn_point=orig_knn_point)
    assert mp.apply(fake) == 2
    pts, feat = T(_data.uniform(4, 3072, 43), cuda), torch.randn(4, 64, 3072, device=cuda)
    got, want = fake.edge_preserve_sampling(feat, pts, 1536, 10), orig_eps(feat, pts, 1536, 10)
    for g_, w_ in zip(got, want):
        assert torch.equal(g_, w_)


def test_model_patches_sa_module_conv_commutes_with_gather(ops, cuda):
    """SA_module.forward with conv2 / conv3 applied before the neighbour gather instead of after
    it (a 1x1 convolution commutes with the gather) against the original order of operations (vrcnet.py:21-57,
    restated): same output and gradients up to the rounding of the convolution algorithm cuDNN picks per shape."""
    import types
    from torch import nn
    from mvp_benchmark_b200 import model_patches as mp
    _, mm = ops

    def get_edge_features(x, idx):                         # model_utils.py:113-124, restated
        batch_size, num_points, k = idx.size()
        idx = (idx + torch.arange(0, batch_size, device=x.device).view(-1, 1, 1) * num_points).view(-1)
        x = x.squeeze(2)
        num_dims = x.size(1)
        x = x.transpose(2, 1).contiguous()
        feature = x.view(batch_size * num_points, -1)[idx, :]
        return feature.view(batch_size, num_points, k, num_dims).permute(0, 3, 2, 1)

    class SA_module(nn.Module):                            # vrcnet.py:21-57, restated
        def __init__(self, in_planes, rel_planes, mid_planes, out_planes, share_planes=8, k=16):
            super().__init__()
            self.share_planes, self.k = share_planes, k
            self.conv1 = nn.Conv2d(in_planes, rel_planes, kernel_size=1)
            self.conv2 = nn.Conv2d(in_planes, rel_planes, kernel_size=1)
            self.conv3 = nn.Conv2d(in_planes, mid_planes, kernel_size=1)
            self.conv_w = nn.Sequential(nn.ReLU(inplace=False),
                                        nn.Conv2d(rel_planes * (k + 1), mid_planes // share_planes, kernel_size=1, bias=False),
                                        nn.ReLU(inplace=False),
                                        nn.Conv2d(mid_planes // share_planes, k * mid_planes // share_planes, kernel_size=1))
            self.activation_fn = nn.ReLU(inplace=False)
            self.conv_out = nn.Conv2d(mid_planes, out_planes, kernel_size=1)

        def forward(self, input):
            x, idx = input
            batch_size, _, _, num_points = x.size()
            identity = x
            x = self.activation_fn(x)
            xn = get_edge_features(x, idx)
            x1, x2, x3 = self.conv1(x), self.conv2(xn), self.conv3(xn)
            x2 = x2.view(batch_size, -1, 1, num_points).contiguous()
            w = self.conv_w(torch.cat([x1, x2], 1)).view(batch_size, -1, self.k, num_points)
            w = w.repeat(1, self.share_planes, 1, 1)
            out = torch.sum(w * x3, dim=2, keepdim=True)
            out = self.conv_out(self.activation_fn(out))
            out = out + identity
            return [out, idx]

    fake = types.ModuleType("fake_vrcnet")
    fake.SA_module, fake.get_edge_features = SA_module, get_edge_features
    SA_module.__module__ = "fake_vrcnet"
    import sys
    sys.modules["fake_vrcnet"] = fake
    try:
        original_forward = SA_module.forward
        torch.manual_seed(0)
        net = SA_module(64, 4, 16, 64, 8, 10).to(cuda)
        B, N = 3, 768
        x0 = torch.randn(B, 64, 1, N, device=cuda)
        idx = torch.randint(0, N, (B, N, 10), device=cuda)
        a = x0.clone().requires_grad_(True)
        want = original_forward(net, [a, idx])[0]
        want.sum().backward()
        gw = {n_: p.grad.clone() for n_, p in net.named_parameters()}
        net.zero_grad()
        assert mp.apply(fake) == 2 and SA_module.forward is mp.sa_module_forward   # get_edge_features and the class
        b = x0.clone().requires_grad_(True)
        got = net([b, idx])[0]
        got.sum().backward()
        scale = want.abs().max().item()
        assert (got - want).abs().max().item() <= 2e-3 * scale
        assert (a.grad - b.grad).abs().max().item() <= 2e-3 * a.grad.abs().max().item()
        for n_, p in net.named_parameters():
            assert (p.grad - gw[n_]).abs().max().item() <= 3e-3 * max(gw[n_].abs().max().item(), 1e-6), n_
    finally:
        del sys.modules["fake_vrcnet"]


def test_model_patches_pointwise_convs(cuda):
    """1x1 convolutions through torch.matmul (model_patches.apply_pointwise_convs): same module parameters, outputs and
    gradients equal to nn.Conv1d / nn.Conv2d up to the rounding of the two libraries' algorithms; other convolutions
    are left alone."""
    from torch import nn
    from mvp_benchmark_b200 import model_patches as mp
    torch.manual_seed(1)
    net = nn.Sequential(nn.Conv2d(64, 4, 1), nn.ReLU(), nn.Conv2d(4, 40, kernel_size=1, bias=False)).to(cuda)
    other = nn.Sequential(nn.Conv2d(8, 8, 3, padding=1), nn.Conv1d(8, 8, 1, groups=2)).to(cuda)
    assert mp.apply_pointwise_convs(other) == 0
    x = torch.randn(5, 64, 3, 700, device=cuda)
    a = x.clone().requires_grad_(True)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                 # the reference in full fp32, like torch.bmm's default
    try:
        want = net(a)
        want.square().sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    ref = {n_: p.grad.clone() for n_, p in net.named_parameters()}
    net.zero_grad()
    assert mp.apply_pointwise_convs(net) == 2
    b = x.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2).clone().requires_grad_(True)   # a non-contiguous view works too
    got = net(b)
    got.square().sum().backward()
    assert got.shape == want.shape and (got - want).abs().max().item() <= 1e-4 * want.abs().max().item()
    assert (a.grad - b.grad).abs().max().item() <= 1e-4 * a.grad.abs().max().item()
    for n_, p in net.named_parameters():
        assert (p.grad - ref[n_]).abs().max().item() <= 1e-3 * ref[n_].abs().max().item(), n_
    big = nn.Conv1d(512, 1024, 1).to(cuda)
    assert mp.apply_pointwise_convs(big, max_weights=16384) == 0   # the caller may leave the wide ones alone
    c1 = nn.Conv1d(16, 3, 1).to(cuda)
    y0 = c1(torch.ones(2, 16, 9, device=cuda))
    assert mp.apply_pointwise_convs(c1) == 1
    torch.testing.assert_close(c1(torch.ones(2, 16, 9, device=cuda)), y0, rtol=1e-3, atol=1e-4)



@pytest.mark.parametrize("B,C,N,k", [(3, 24, 700, 16), (2, 5, 33, 32), (1, 96, 2048, 20), (2, 8, 129, 1)])
def test_topk_rows_sqdist_is_the_original_formula(cuda, B, C, N, k):
    """fused.topk_rows_sqdist (mvp_topk_rows_sqdist) against topk_rows / torch.topk on the score matrix the original
    builds (completion/model_utils.py:243-245): the on-the-fly score is torch's, bit for bit, so values and indices are
    identical — including on a lattice of features, where exact ties abound."""
    from mvp_benchmark_b200 import fused
    g = torch.Generator(device=cuda).manual_seed(N + k)
    for lattice in (False, True):
        x = torch.randn(B, C, N, device=cuda, generator=g)
        if lattice:
            x = torch.round(x)
        inner = -2 * torch.matmul(x.transpose(2, 1).contiguous(), x)
        xx = torch.sum(x ** 2, dim=1, keepdim=True)
        score = -xx - inner - xx.transpose(2, 1).contiguous()
        wv, wi = fused.topk_rows(score, k)
        v, i = fused.topk_rows_sqdist(torch.matmul(x.transpose(2, 1).contiguous(), x), torch.sum(x ** 2, dim=1), k)
        assert torch.equal(v, wv) and torch.equal(i, wi)
        if not lattice:
            tv, ti = score.topk(k, dim=-1)
            assert torch.equal(v, tv)


def test_model_patches_pcn_decoder(cuda):
    """PCN_decoder.forward (completion/models/pcn.py:48-71, restated) against model_patches.pcn_decoder_forward: the
    global feature's share of conv1 as a per-cloud vector — same outputs and parameter gradients up to TF32 / summation
    order; fused.add_per_cloud against the broadcasting add it stands for."""
    import types
    import torch.nn.functional as F
    from torch import nn
    from mvp_benchmark_b200 import fused, model_patches as mp

    class PCN_decoder(nn.Module):
        def __init__(self, num_coarse, num_fine, scale, cat_feature_num):
            super().__init__()
            self.num_coarse, self.num_fine, self.scale = num_coarse, num_fine, scale
            self.fc1, self.fc2, self.fc3 = nn.Linear(1024, 1024), nn.Linear(1024, 1024), nn.Linear(1024, num_coarse * 3)
            r = int(scale ** 0.5)
            gx, gy = torch.meshgrid(torch.linspace(-0.05, 0.05, r), torch.linspace(-0.05, 0.05, r), indexing="ij")
            self.grid = torch.stack((gx, gy), -1).view(-1, 2).transpose(0, 1).contiguous().cuda()
            self.conv1, self.conv2, self.conv3 = nn.Conv1d(cat_feature_num, 512, 1), nn.Conv1d(512, 512, 1), nn.Conv1d(512, 3, 1)

        def forward(self, x):
            batch_size = x.size()[0]
            coarse = F.relu(self.fc1(x))
            coarse = F.relu(self.fc2(coarse))
            coarse = self.fc3(coarse).view(-1, 3, self.num_coarse)
            grid = self.grid.clone().detach()
            grid_feat = grid.unsqueeze(0).repeat(batch_size, 1, self.num_coarse).contiguous().cuda()
            point_feat = ((coarse.transpose(1, 2).contiguous()).unsqueeze(2).repeat(1, 1, self.scale, 1).view(
                -1, self.num_fine, 3)).transpose(1, 2).contiguous()
            global_feat = x.unsqueeze(2).repeat(1, 1, self.num_fine)
            feat = torch.cat((grid_feat, point_feat, global_feat), 1)
            center = ((coarse.transpose(1, 2).contiguous()).unsqueeze(2).repeat(1, 1, self.scale, 1).view(
                -1, self.num_fine, 3)).transpose(1, 2).contiguous()
            fine = self.conv3(F.relu(self.conv2(F.relu(self.conv1(feat))))) + center
            return coarse, fine

    torch.manual_seed(3)
    dec = PCN_decoder(96, 96 * 4, 4, 1024 + 5).to(cuda)
    x = torch.randn(3, 1024, device=cuda)
    original = PCN_decoder.forward
    c0, f0 = dec(x)
    (f0.square().sum() + c0.sum()).backward()
    ref = {n_: p.grad.clone() for n_, p in dec.named_parameters()}
    dec.zero_grad()
    fake = types.SimpleNamespace(PCN_decoder=PCN_decoder)
    try:
        assert mp.apply(fake) == 1 and PCN_decoder.forward is mp.pcn_decoder_forward
        c1, f1 = dec(x)
        (f1.square().sum() + c1.sum()).backward()
    finally:
        PCN_decoder.forward = original
    assert torch.equal(c0, c1) and (f1 - f0).abs().max().item() <= 3e-3 * f0.abs().max().item()
    for n_, p in dec.named_parameters():      # TF32 sums in another order, a few ReLU signs near zero: compare in norm
        assert (p.grad - ref[n_]).norm().item() <= 2e-2 * max(ref[n_].norm().item(), 1e-6), n_
    y = torch.randn(4, 7, 130, device=cuda)
    v = torch.randn(4, 7, device=cuda)
    a, b = y.clone().requires_grad_(True), v.clone().requires_grad_(True)
    out = fused.add_per_cloud(a * 1.0, b, relu=True)
    want = F.relu(y + v[:, :, None])
    assert torch.equal(out, want)
    go = torch.randn_like(out)
    out.backward(go)
    assert torch.equal(a.grad, go * (want > 0))
    assert (b.grad - (go * (want > 0)).sum(2)).abs().max().item() <= 1e-5 * go.abs().sum(2).max().item()


def test_fscore_and_emd_loss_epilogues(cuda):
    """fused.fscore (mvp_fscore) against the oracle and the reference's torch lines (utils/metrics/CD/fscore.py:12-15):
    identical, NaN -> 0 included; fused.emd_loss against `torch.sqrt(dist).mean(1)` (model_utils.py:84) forward and
    backward at 1e-5; model_patches.calc_emd returns what the original returns."""
    import types
    import oracle
    from mvp_benchmark_b200 import fused, model_patches as mp
    rng = np.random.default_rng(11)
    d1 = (rng.random((5, 2048), dtype=np.float32) * 3e-4).astype(np.float32)
    d2 = (rng.random((5, 777), dtype=np.float32) * 3e-4).astype(np.float32)
    d1[2] += 1.0
    d2[2] += 1.0
    t1, t2 = torch.from_numpy(d1).to(cuda), torch.from_numpy(d2).to(cuda)
    f, p1, p2 = fused.fscore(t1, t2)
    of, o1, o2 = oracle.fscore(d1, d2)
    assert np.array_equal(p1.cpu().numpy(), o1) and np.array_equal(p2.cpu().numpy(), o2) and np.array_equal(f.cpu().numpy(), of)
    w1 = torch.mean((t1 < 0.0001).float(), dim=1)
    w2 = torch.mean((t2 < 0.0001).float(), dim=1)
    wf = 2 * w1 * w2 / (w1 + w2)
    wf[torch.isnan(wf)] = 0
    torch.testing.assert_close(p1, w1, rtol=1e-6, atol=0), torch.testing.assert_close(f, wf, rtol=1e-6, atol=0)
    a = (t1 + 1e-6).clone().requires_grad_(True)
    b = (t1 + 1e-6).clone().requires_grad_(True)
    got, want = fused.emd_loss(a), torch.sqrt(b).mean(1)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=0)
    go = torch.randn(5, device=cuda)
    got.backward(go), want.backward(go)
    torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=0)

    def original(output, gt, eps=0.005, iterations=50):    # model_utils.py:80-85, restated
        import metrics
        dist, _ = metrics.emd()(output, gt, eps, iterations)
        return torch.sqrt(dist).mean(1)

    import mvp_benchmark_b200
    mvp_benchmark_b200.install()
    fake = types.SimpleNamespace(calc_emd=original)
    assert mp.apply(fake) == 1
    x = torch.rand(2, 1024, 3, device=cuda)
    y = torch.rand(2, 1024, 3, device=cuda)
    torch.testing.assert_close(fake.calc_emd(x, y), original(x, y), rtol=1e-5, atol=0)


def _tf32(t):
    """round to nearest (ties away) to TF32's 10-bit mantissa, as cvt.rna.tf32.f32 does"""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("B,C,O,N,relu,bias", [(2, 32, 16, 128, False, False), (2, 64, 256, 384, False, True), (3, 3, 5, 77, True, True),
                                               (2, 67, 300, 130, False, True), (2, 512, 1024, 256, True, True),
                                               (1, 130, 16, 1000, False, False), (2, 8, 4, 2048, False, True),
                                               (5, 128, 256, 700, True, True), (2, 256, 256, 130, False, True), (0, 8, 8, 16, False, True)])
def test_pointwise_conv_tensor_core(cuda, B, C, O, N, relu, bias):
    """mvp_pointwise_conv (tcgen05.mma kind::tf32, csrc/pointwise.cu; resident-weight and streaming variants by shape)
    against the same contraction in fp64 of the TF32-rounded operands: what is left is fp32 accumulation order, so the
    tolerance is 1e-4 of the output scale (against the unrounded operands: TF32's 2^-11 per operand)."""
    from mvp_benchmark_b200 import fused
    g = torch.Generator(device=cuda).manual_seed(B * 1000 + C + O + N)
    x = torch.randn(B, C, N, device=cuda, generator=g)
    w = torch.randn(O, C, device=cuda, generator=g) / C ** 0.5
    bs = torch.randn(O, device=cuda, generator=g) if bias else None
    y = fused.pointwise_conv(x, w, bs, relu)
    assert y.shape == (B, O, N)
    if B == 0:
        return
    ref = torch.matmul(_tf32(w).double(), _tf32(x).double())
    full = torch.matmul(w.double(), x.double())
    if bias:
        ref, full = ref + bs.double().view(1, -1, 1), full + bs.double().view(1, -1, 1)
    if relu:
        ref, full = ref.clamp_min(0), full.clamp_min(0)
    scale = full.abs().max().item()
    assert (y.double() - ref).abs().max().item() <= 1e-4 * scale
    assert (y.double() - full).abs().max().item() <= 2e-3 * scale


def test_pointwise_conv_exact_cases_and_gradients(cuda):
    """One-hot operands land where they should (the shared-memory operand layout: every (point, channel) pair of a tile
    boundary case), values representable in TF32 come out exact, and the autograd Function's gradients agree with
    nn.Conv1d's; the masked entry (input gradient behind a ReLU) equals the unmasked one on the masked gradient."""
    import torch.nn.functional as F
    from mvp_benchmark_b200 import fused
    for (C, O, N, k, co, p) in [(32, 16, 128, 0, 0, 0), (32, 16, 128, 5, 3, 77), (64, 256, 256, 37, 200, 130),
                                (8, 16, 128, 7, 15, 127), (40, 48, 300, 33, 47, 299), (300, 70, 129, 299, 69, 128)]:
        x = torch.zeros(1, C, N, device=cuda); x[0, k, p] = 1.0
        w = torch.zeros(O, C, device=cuda); w[co, k] = 2.0
        y = fused.pointwise_conv(x, w)
        nz = torch.nonzero(y)
        assert nz.shape[0] == 1 and nz[0].tolist() == [0, co, p] and float(y[0, co, p]) == 2.0
    g = torch.Generator(device=cuda).manual_seed(3)
    xi = torch.randint(-8, 9, (3, 96, 515), device=cuda, generator=g).float()
    wi = torch.randint(-4, 5, (80, 96), device=cuda, generator=g).float()
    assert torch.equal(fused.pointwise_conv(xi, wi), torch.matmul(wi, xi))        # small integers: exact in TF32 and fp32
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for (B, C, O, N) in [(4, 64, 256, 512), (2, 130, 48, 300), (2, 512, 640, 256), (3, 128, 256, 700)]:
            x = torch.randn(B, C, N, device=cuda, generator=g).requires_grad_(True)
            w = (torch.randn(O, C, 1, device=cuda, generator=g) / C ** 0.5).requires_grad_(True)
            bs = torch.randn(O, device=cuda, generator=g).requires_grad_(True)
            go = torch.randn(B, O, N, device=cuda, generator=g)
            for relu in (False, True):
                y = fused.pointwise_conv(x, w, bs, relu); y.backward(go)
                got = [y.detach(), x.grad.clone(), w.grad.clone(), bs.grad.clone()]
                x.grad = w.grad = bs.grad = None
                y2 = F.conv1d(x, w, bs)
                if relu:  # the sign pattern of THIS output: a pre-activation within TF32 rounding of zero may fall either way
                    y2 = y2 * (got[0] > 0)
                y2.backward(go)
                for a, r in zip(got, [y2.detach(), x.grad, w.grad, bs.grad]):
                    assert (a - r).abs().max().item() <= 3e-3 * r.abs().max().item()
                x.grad = w.grad = bs.grad = None
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    go = torch.randn(2, 48, 300, device=cuda, generator=g); yy = torch.randn(2, 48, 300, device=cuda, generator=g)
    wt = torch.randn(130, 48, device=cuda, generator=g)
    m1, m2 = fused._pointwise_conv_raw(go, wt, None, False, mask=yy), fused._pointwise_conv_raw(go * (yy > 0), wt, None)
    assert (m1 - m2).abs().max().item() <= 1e-5 * m2.abs().max().item()   # two kernels (streaming / resident), same products


@pytest.mark.parametrize("B,C,O,N,bias", [(2, 32, 16, 128, True), (3, 64, 256, 384, True), (3, 3, 5, 77, True), (2, 67, 300, 130, False),
                                          (4, 255, 40, 515, True), (2, 256, 128, 1000, False), (16, 64, 4, 3072, True), (1, 8, 8, 4, True),
                                          (5, 128, 256, 700, True)])
def test_pointwise_wgrad_tensor_core(cuda, B, C, O, N, bias):
    """mvp_pointwise_wgrad (tcgen05, K = points, a row of ones for the bias gradient) against fp64 sums of the
    TF32-rounded operands (1e-4 of the scale: accumulation order) and of the raw operands (TF32's rounding);
    deterministic: two runs are bit-identical."""
    from mvp_benchmark_b200 import fused
    g_ = torch.Generator(device=cuda).manual_seed(C * 7 + O)
    x = torch.randn(B, C, N, device=cuda, generator=g_)
    g = torch.randn(B, O, N, device=cuda, generator=g_)
    gw, gb = fused._pointwise_wgrad_raw(g, x, bias)
    gw2, gb2 = fused._pointwise_wgrad_raw(g, x, bias)
    assert torch.equal(gw, gw2) and (not bias or torch.equal(gb, gb2))
    ref = torch.einsum("bon,bcn->oc", _tf32(g).double(), _tf32(x).double())
    full = torch.einsum("bon,bcn->oc", g.double(), x.double())
    scale = full.abs().max().item()
    assert (gw.double() - ref).abs().max().item() <= 1e-4 * scale
    assert (gw.double() - full).abs().max().item() <= 3e-3 * scale
    if bias:
        sref = _tf32(g).double().sum((0, 2))
        assert (gb.double() - sref).abs().max().item() <= 1e-4 * g.abs().double().sum((0, 2)).max().item()
    else:
        assert gb is None


@pytest.mark.parametrize("shape", [(4, 24, 700, 16), (3, 5, 77, 20), (2, 9, 31), (5, 1), (2, 3, 50, 7), (0, 4, 8)])
def test_max_last_vs_torch(cuda, shape):
    """fused.max_last (mvp_max_last / _grad) against torch.max(x, -1): values bit-identical, the FIRST position of the
    maximum among equals (torch's choice, which routes the gradient), NaN wins (the first one), gradient identical."""
    from mvp_benchmark_b200 import fused
    g = torch.Generator(device=cuda).manual_seed(len(shape) * 31 + shape[-1])
    x = (torch.round(torch.randn(*shape, device=cuda, generator=g) * 2) / 2).requires_grad_(True)   # many exact ties
    v, a = fused.max_last(x)
    wv, wi = torch.max(x.detach(), -1)
    assert v.shape == wv.shape and torch.equal(v, wv)
    if x.numel():
        first = (x.detach() == wv.unsqueeze(-1)).float().argmax(-1)
        assert torch.equal(a.long(), first)
    go = torch.randn(*shape[:-1], device=cuda, generator=g)
    v.backward(go)
    want = torch.zeros_like(x).scatter_(-1, a.long().unsqueeze(-1), go.unsqueeze(-1)) if x.numel() else torch.zeros_like(x)
    assert torch.equal(x.grad, want)
    if x.numel() > 8:
        y = x.detach().clone()
        y.view(-1)[3] = float("nan")
        v2, a2 = fused.max_last(y)
        wv2, _ = torch.max(y, -1)
        assert torch.equal(torch.isnan(v2), torch.isnan(wv2)) and torch.equal(torch.nan_to_num(v2), torch.nan_to_num(wv2))
        assert int(a2.view(-1)[3 // shape[-1]]) == 3 % shape[-1]


@pytest.mark.parametrize("B,C,N", [(8, 1024, 2048), (3, 5, 77), (64, 4, 3072), (2, 1000, 6), (7, 33, 1)])
def test_bias_add_and_channel_sum(cuda, B, C, N):
    """mvp_bias_add: bit-identical to torch's broadcasting add (one IEEE add per element); mvp_channel_sum: the
    per-channel sum over clouds and points within fp32 accumulation error of the fp64 sum, and deterministic."""
    from mvp_benchmark_b200 import fused
    g = torch.Generator(device=cuda).manual_seed(C)
    y = torch.randn(B, C, N, device=cuda, generator=g); bs = torch.randn(C, device=cuda, generator=g)
    assert torch.equal(fused.bias_add_(y.clone(), bs), y + bs.view(1, -1, 1))
    assert torch.equal(fused.bias_add_(y.clone(), bs, relu=True), (y + bs.view(1, -1, 1)).clamp_min(0))
    s1, s2 = fused.channel_sum(y), fused.channel_sum(y)
    assert torch.equal(s1, s2)
    want = y.double().sum((0, 2))
    assert (s1.double() - want).abs().max().item() <= 2e-6 * y.abs().double().sum((0, 2)).max().item()
    y4 = y.view(B, C, N, 1)
    assert torch.equal(fused.channel_sum(y4), s1)


def test_model_patches_conv_routes(cuda):
    """_pointwise_conv_forward's three routes give what nn.Conv gives (TF32 tolerance) with gradients: the tcgen05 kernel
    (64 -> 256 over a long tensor), the library GEMM with this repository's bias kernels (512 -> 640), cuDNN for the rest."""
    from torch import nn
    from mvp_benchmark_b200 import model_patches as mp
    torch.manual_seed(5)
    for conv, x in ((nn.Conv2d(64, 256, 1), torch.randn(16, 64, 1, 1100, device=cuda)),
                    (nn.Conv1d(512, 640, 1), torch.randn(16, 512, 600, device=cuda)),
                    (nn.Conv1d(300, 20, 1), torch.randn(4, 300, 50, device=cuda))):
        conv = conv.to(cuda)
        a = x.clone().requires_grad_(True)
        want = conv(a); want.square().sum().backward()
        ref = [a.grad.clone()] + [p.grad.clone() for p in conv.parameters()]
        conv.zero_grad()
        assert mp.apply_pointwise_convs(conv) == 1
        b = x.clone().requires_grad_(True)
        got = conv(b); got.square().sum().backward()
        assert (got - want).abs().max().item() <= 3e-3 * want.abs().max().item()
        for u, v in zip([b.grad] + [p.grad for p in conv.parameters()], ref):
            assert (u - v).abs().max().item() <= 5e-3 * v.abs().max().item()


@pytest.mark.parametrize("shape,k", [((3, 700, 700), 16), ((2, 5, 33), 32), ((4096, 40), 1), ((2, 3, 2048, 2048), 20), ((1, 9, 31), 31)])
def test_topk_rows_vs_torch(cuda, shape, k):
    """fused.topk_rows (mvp_topk_rows) against torch.topk: identical values; indices identical where the values are
    distinct, ascending index inside runs of equal values (lattice-like scores), NaN never selected."""
    from mvp_benchmark_b200 import fused
    g = torch.Generator(device=cuda).manual_seed(9)
    x = torch.randn(*shape, device=cuda, generator=g)
    v, i = fused.topk_rows(x, k)
    wv, wi = x.topk(k, dim=-1)
    assert i.dtype == torch.int64 and torch.equal(v, wv) and torch.equal(i, wi)
    q = torch.round(x * 2) / 2                               # many exact ties
    q[..., 0] = float("nan")
    v, i = fused.topk_rows(q, min(k, shape[-1] - 1))
    wv, _ = torch.nan_to_num(q, nan=float("-inf")).topk(min(k, shape[-1] - 1), dim=-1)
    assert torch.equal(v, wv) and torch.equal(torch.gather(q, -1, i), v) and (i != 0).all()
    same = v[..., 1:] == v[..., :-1]
    assert (i[..., 1:][same] > i[..., :-1][same]).all()      # ties in ascending index
    srt = i.sort(dim=-1)[0]
    assert (srt[..., 1:] != srt[..., :-1]).all()             # no index twice



@pytest.mark.parametrize("B,C,Cw,N,K", [(3, 16, 2, 3072, 20), (2, 128, 16, 384, 10), (2, 24, 3, 777, 5), (1, 8, 8, 100, 1)])
def test_neighbor_weighted_sum_vs_torch(ops, cuda, B, C, Cw, N, K):
    """fused.neighbor_weighted_sum (SA_module's aggregation, vrcnet.py:49-52) against the torch sequence it replaces —
    gather of the neighbours' features, repeat of the weights over share_planes, multiply, sum over k — forward and
    both gradients."""
    from mvp_benchmark_b200 import fused
    _, mm = ops
    g_ = torch.Generator(device=cuda).manual_seed(13)
    y0 = torch.randn(B, C, N, device=cuda, generator=g_)
    w0 = torch.randn(B, Cw, K, N, device=cuda, generator=g_)
    idx = torch.randint(0, N, (B, N, K), device=cuda, generator=g_, dtype=torch.int32)
    ya, wa = y0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    yb, wb = y0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    got = fused.neighbor_weighted_sum(ya, idx, wa)
    x3 = mm.grouping_operation(yb, idx.transpose(1, 2).contiguous())          # (B, C, K, N), as get_edge_features
    want = torch.sum(wb.repeat(1, C // Cw, 1, 1) * x3, dim=2)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    go = torch.randn_like(got)
    got.backward(go), want.backward(go)
    torch.testing.assert_close(wa.grad, wb.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ya.grad, yb.grad, rtol=1e-4, atol=1e-4)
