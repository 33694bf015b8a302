"""CPU-only: the Python layer mirrors the reference's package surface and fails loudly off-GPU."""
import os

import pytest
import torch


def test_package_surface_matches_reference_names(ops):
    metrics, mm3d = ops
    assert metrics.__all__ == ['cd', 'fscore', 'emd']                      # utils/metrics/__init__.py:4-6
    for name in ['ball_query', 'knn', 'furthest_point_sample', 'furthest_point_sample_with_dist',
                 'three_interpolate', 'three_nn', 'gather_points', 'grouping_operation', 'group_points',
                 'GroupAll', 'QueryAndGroup', 'Points_Sampler', 'get_compiler_version',
                 'get_compiling_cuda_version']:                              # utils/mm3d_pn2/__init__.py:7-20
        assert hasattr(mm3d, name), name
    import types
    assert isinstance(mm3d.group_points, types.ModuleType)                  # resolves to the submodule (ops/__init__.py:9)
    assert mm3d.group_points.grouping_operation is mm3d.grouping_operation
    from metrics.CD.chamfer3D.dist_chamfer_3D import chamfer_3DDist, chamfer_3DFunction  # noqa: F401
    from metrics.EMD.emd_module import emdFunction, emdModule  # noqa: F401
    assert metrics.cd is chamfer_3DDist and metrics.emd is emdModule


def test_reference_import_lines_resolve(ops):
    # completion/model_utils.py:20-21, completion/models/vrcnet.py:17-18
    from metrics import cd, fscore, emd  # noqa: F401
    from mm3d_pn2 import furthest_point_sample, gather_points, grouping_operation, ball_query, three_nn  # noqa: F401
    from mm3d_pn2 import three_interpolate  # noqa: F401


def test_no_cpu_fallback(ops):
    metrics, mm3d = ops
    from mvp_benchmark_b200._lib import MvpOpsError
    a = torch.rand(2, 16, 3)
    with pytest.raises(MvpOpsError):
        metrics.cd()(a, a)
    with pytest.raises(MvpOpsError):
        mm3d.furthest_point_sample(a, 4)
    with pytest.raises(MvpOpsError):
        mm3d.three_nn(a, a)
    with pytest.raises(MvpOpsError):
        mm3d.gather_points(torch.rand(2, 4, 16), torch.zeros(2, 3, dtype=torch.int32))


def test_missing_extension_fails_loudly(tmp_path, monkeypatch):
    import importlib.util
    from mvp_benchmark_b200 import _lib
    spec = importlib.util.spec_from_file_location("_lib_copy", _lib.__file__.replace("_lib.py", "_lib.py"))
    mod = importlib.util.module_from_spec(spec)
    src = open(_lib.__file__).read().replace('os.path.join(_HERE, "libmvp_ops.so")', repr(str(tmp_path / "nope.so")))
    with pytest.raises(ImportError, match="no CPU fallback"):
        exec(compile(src, "_lib_copy.py", "exec"), mod.__dict__)


def test_fscore_contract(ops):
    metrics, _ = ops
    d1 = torch.tensor([[0.0, 1e-5, 1.0, 2.0], [1.0, 1.0, 1.0, 1.0]])
    d2 = torch.tensor([[0.0, 0.0], [1.0, 1.0]])
    f, p1, p2 = metrics.fscore(d1, d2)
    assert torch.allclose(p1, torch.tensor([0.5, 0.0])) and torch.allclose(p2, torch.tensor([1.0, 0.0]))
    assert torch.allclose(f, torch.tensor([2 * 0.5 / 1.5, 0.0]))            # NaN -> 0 (fscore.py:15)


def test_chamfer_python_matches_reference_python_golden(ops):
    import numpy as np
    from metrics.CD import chamfer_python
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "chamfer_python_ref.npz"))
    d1, d2, i1, i2 = chamfer_python.distChamfer(torch.from_numpy(G["unit_xyz1"]), torch.from_numpy(G["unit_xyz2"]))
    assert np.allclose(d1.numpy(), G["unit_dist1"], atol=1e-7) and (i1.numpy() == G["unit_idx1"]).all()
    assert np.allclose(d2.numpy(), G["unit_dist2"], atol=1e-7) and (i2.numpy() == G["unit_idx2"]).all()
    assert d1.dtype == torch.float32 and i1.dtype == torch.int32


def test_calc_square_dist_and_sampler_config(ops):
    _, mm3d = ops
    from mm3d_pn2.ops.furthest_point_sample.utils import calc_square_dist
    a, b = torch.rand(2, 5, 4), torch.rand(2, 7, 4)
    ref = ((a[:, :, None] - b[:, None]) ** 2).sum(-1)
    assert torch.allclose(calc_square_dist(a, b, norm=False), ref, atol=1e-5)
    assert torch.allclose(calc_square_dist(a, b, norm=True), ref.clamp_min(0).sqrt() / 4, atol=1e-4)
    with pytest.raises(ValueError):
        mm3d.Points_Sampler([4], ['X-FPS'])
    s = mm3d.Points_Sampler([4, 4], ['D-FPS', 'F-FPS'], [8, -1])
    assert len(s.samplers) == 2
    g = mm3d.GroupAll(use_xyz=True)(torch.rand(2, 6, 3), None, torch.rand(2, 5, 6))
    assert g.shape == (2, 8, 1, 6)


def test_model_patches_fall_through_for_what_the_fused_ops_do_not_cover():
    """Opt-in model patches: CPU tensors, feature-space kNN (C != 3) and k beyond the fused op's
    range go to the ORIGINAL function of the patched module — the patches never compute on the CPU themselves."""
    import types
    import torch
    from mvp_benchmark_b200 import model_patches as mp
    seen = []
    fake = types.SimpleNamespace(
        knn=lambda x, k: seen.append(("knn", tuple(x.shape), k)) or "orig-knn",
        knn_point=lambda pk, a, b: seen.append(("knn_point", pk)) or ("orig-d", "orig-i"),
        get_edge_features=lambda x, idx: seen.append(("gef", tuple(x.shape))) or "orig-gef")
    assert mp.apply(fake) == 3
    assert fake.knn(torch.zeros(2, 3, 50), 4) == "orig-knn"                  # CPU tensor
    assert fake.knn(torch.zeros(2, 64, 50), 4) == "orig-knn"                 # feature space
    assert fake.knn_point(5, torch.zeros(2, 50, 3), torch.zeros(2, 10, 3)) == ("orig-d", "orig-i")
    assert fake.get_edge_features(torch.zeros(2, 8, 1, 50), torch.zeros(2, 50, 4, dtype=torch.long)) == "orig-gef"
    assert [s[0] for s in seen] == ["knn", "knn", "knn_point", "gef"]


def test_model_patches_layer_routes_fall_through_on_the_cpu():
    """The class- and module-level patches of round 2 (1x1 layers, PCN decoder, ECG's Dense_conv / get_graph_feature):
    on CPU tensors every one of them runs the ORIGINAL code — nothing of this repository computes on the CPU."""
    import types
    import torch
    from torch import nn
    from mvp_benchmark_b200 import model_patches as mp
    torch.manual_seed(0)
    net = nn.Sequential(nn.Conv1d(64, 256, 1), nn.ReLU(), nn.Conv1d(256, 4, 1), nn.Conv2d(4, 4, 3, padding=1) if False else nn.Identity())
    x = torch.randn(2, 64, 50)
    want = net(x)
    assert mp.apply_pointwise_convs(net) == 2 and mp.apply_pointwise_convs(net, max_weights=16) == 0
    assert torch.equal(net(x), want)                        # module's own forward: not CUDA
    calls = []

    class PCN_decoder(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv1d(1029, 8, 1)

        def forward(self, x):
            calls.append("pcn")
            return "coarse", "fine"

    fake = types.SimpleNamespace(PCN_decoder=PCN_decoder,
                                 get_graph_feature=lambda x, k=20, minus_center=True: calls.append("ggf") or "orig-ggf")
    assert mp.apply(fake) == 2
    try:
        assert PCN_decoder()(torch.zeros(2, 1024)) == ("coarse", "fine")
        assert fake.get_graph_feature(torch.zeros(2, 8, 30), 4) == "orig-ggf"
    finally:
        PCN_decoder.forward = PCN_decoder._mvp_original_forward
    assert calls == ["pcn", "ggf"]
