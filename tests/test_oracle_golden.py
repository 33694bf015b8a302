"""CPU-only: pins the oracle (oracle/oracle.c) to
  (1) the reference's own pure-torch Chamfer, with the reference's own assertions
      (utils/metrics/CD/unit_test.py:22-33: summed mean-squared error < 1e-8, indices identical);
  (2) outputs of the reference's CUDA kernels recompiled for sm_100a and run on a B200
      (tests/golden/ref_cuda_golden.npz, made by tests/golden/make_golden_gpu.py) — every op.
"""
import os

import numpy as np
import pytest

import _cases
from _impls import OracleImpl

HERE = os.path.dirname(os.path.abspath(__file__))
PY_GOLD = os.path.join(HERE, "golden", "chamfer_python_ref.npz")
CUDA_GOLD = os.path.join(HERE, "golden", "ref_cuda_golden.npz")


@pytest.fixture(scope="module")
def impl():
    return OracleImpl()


@pytest.mark.parametrize("case", ["unit", "unit1", "c1"])
def test_oracle_chamfer_vs_reference_python(impl, case):
    G = np.load(PY_GOLD)
    d1, d2, i1, i2 = impl.chamfer_forward(G[case + "_xyz1"], G[case + "_xyz2"])
    err = ((d1 - G[case + "_dist1"]) ** 2).mean() + ((d2 - G[case + "_dist2"]) ** 2).mean()
    assert err < 1e-8                                   # unit_test.py:23-27
    assert (i1 == G[case + "_idx1"]).all() and (i2 == G[case + "_idx2"]).all()  # unit_test.py:29-33


def _cuda_cases():
    if not os.path.isfile(CUDA_GOLD):
        return []
    return [nm for nm, _ in _cases.all_cases(np.load(CUDA_GOLD))]


@pytest.mark.skipif(not os.path.isfile(CUDA_GOLD), reason="ref_cuda_golden.npz not generated yet")
@pytest.mark.parametrize("case", _cuda_cases())
def test_oracle_vs_reference_cuda_golden(impl, case):
    G = dict(np.load(CUDA_GOLD))
    fn = dict(_cases.all_cases(G))[case]
    fn(impl, _cases.group(G, case), case)


def test_oracle_fps_block_size_rule():
    import oracle
    # furthest_point_sample_cuda.cu:11-15 — largest power of two <= n, capped at 1024
    for n, t in [(1, 1), (2, 2), (3, 2), (100, 64), (384, 256), (768, 512), (1024, 1024), (1536, 1024), (16384, 1024)]:
        assert oracle.fps_block_size(n) == t


def test_oracle_emd_rejects_bad_sizes():
    import oracle
    x = np.zeros((1, 1000, 3), np.float32)
    with pytest.raises(ValueError):
        oracle.emd_forward(x, x, 0.005, 5)     # n % 1024 != 0, emd_cuda.cu:246-249


def test_oracle_emd_invariants():
    import oracle
    import _data
    x1, x2 = _data.uniform(2, 1024, 5), _data.uniform(2, 1024, 6)
    d, a = oracle.emd_forward(x1, x2, 0.005, 50)
    _cases.emd_consistent(x1, x2, d, a)
    assert len(np.unique(a[0])) > 900          # near-bijection after 50 rounds (emd_module.py:99)
    d3k, a3k = oracle.emd_forward(x1, x2, 0.002, 3000)
    assert len(np.unique(a3k[0])) >= len(np.unique(a[0]))


KNN_GOLD = os.path.join(HERE, "golden", "knn_model_utils.npz")


@pytest.mark.parametrize("case", ["self16", "self20", "point10", "point2_small", "point4_sphere"])
def test_oracle_knn_points_vs_reference_model_utils(case):
    """oracle.knn_points against the reference's own torch `knn` / `knn_point` (completion/model_utils.py:242-259,
    outputs stored by tests/golden/make_golden_knn.py).  The reference ranks by the matmul expansion
    -|x|^2 + 2x.y - |y|^2 (rounding ~1e-6 at unit scale): wherever its k-th and (k+1)-th distances are further
    apart than that, the neighbour SET must be identical; everywhere, the exact distances of the reference's
    choice and ours must agree to that rounding, rank by rank."""
    import oracle
    G = np.load(KNN_GOLD)
    cloud, queries, ref_idx, k = G[case + "_cloud"], G[case + "_queries"], G[case + "_idx"], int(G[case + "_k"])
    d, idx = oracle.knn_points(k, cloud, queries)
    assert idx.shape == ref_idx.shape and (np.diff(d, axis=2) >= 0).all()
    clear = G[case + "_kth_gap"] > 5e-6
    assert clear.mean() > 0.9
    same_set = (np.sort(idx, axis=2) == np.sort(ref_idx, axis=2)).all(axis=2)
    assert same_set[clear].all()
    b = np.arange(cloud.shape[0])[:, None, None]
    exact_ref = ((queries[:, :, None, :].astype(np.float64) - cloud[b, ref_idx].astype(np.float64)) ** 2).sum(-1)
    np.testing.assert_allclose(d, exact_ref, atol=5e-6, rtol=0)
    if case + "_negdist" in G.files:           # knn_point also returns -distance (model_utils.py:258)
        np.testing.assert_allclose(-d, G[case + "_negdist"], atol=5e-6, rtol=0)


@pytest.mark.parametrize("case", ["vrcnet", "square", "tiny"])
def test_oracle_chamfer_loss_vs_reference_calc_cd(case):
    """oracle.chamfer_loss against the reference's own calc_cd (completion/model_utils.py:67-77, run on CPU over its
    pure-torch Chamfer by tests/golden/make_golden_loss.py): 1e-5 relative, north_star's floating-point bar."""
    import oracle
    G = np.load(os.path.join(HERE, "golden", "calc_cd_epilogue.npz"))
    cd_p, cd_t = oracle.chamfer_loss(G[case + "_dist1"], G[case + "_dist2"])
    np.testing.assert_allclose(cd_p, G[case + "_cd_p"], rtol=1e-5, atol=0)
    np.testing.assert_allclose(cd_t, G[case + "_cd_t"], rtol=1e-5, atol=0)


@pytest.mark.parametrize("B,C,O,N", [(2, 3, 5, 7), (3, 16, 8, 33), (1, 67, 30, 12)])
def test_oracle_pointwise_conv_vs_torch_layers(B, C, O, N):
    """oracle.pointwise_conv / _grads against what the reference's models call — torch's own nn.Conv1d(kernel_size=1)
    forward and autograd on the CPU (fp32): 1e-5 relative, north_star's floating-point bar; max_last against torch.max
    (values and index identical, ties and NaN included)."""
    import oracle
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(B * 100 + C)
    x = rng.standard_normal((B, C, N)).astype(np.float32)
    w = rng.standard_normal((O, C)).astype(np.float32)
    bs = rng.standard_normal(O).astype(np.float32)
    g = rng.standard_normal((B, O, N)).astype(np.float32)
    tx, tw, tb = (torch.from_numpy(a).requires_grad_(True) for a in (x, w[:, :, None].copy(), bs))
    ty = F.conv1d(tx, tw, tb)
    ty.backward(torch.from_numpy(g))
    np.testing.assert_allclose(oracle.pointwise_conv(x, w, bs), ty.detach().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle.pointwise_conv(x, w, bs, relu=True), F.relu(ty).detach().numpy(), rtol=1e-5, atol=1e-5)
    gx, gw, gb = oracle.pointwise_conv_grads(x, w, g)
    np.testing.assert_allclose(gx, tx.grad.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gw, tw.grad.numpy()[:, :, 0], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(gb, tb.grad.numpy(), rtol=1e-5, atol=1e-4)
    q = np.round(x * 2) / 2
    q[0, 0, 1] = np.nan
    v, a = oracle.max_last(q)
    tv, ta = torch.max(torch.from_numpy(q), -1)
    np.testing.assert_array_equal(v, tv.numpy())
    np.testing.assert_array_equal(a, ta.numpy())


def test_oracle_fscore_vs_reference_fscore():
    """oracle.fscore against the reference's own fscore (utils/metrics/CD/fscore.py, restated line by line and run by
    torch on the CPU): identical, including a cloud with no point under the threshold (NaN -> 0) and odd sizes."""
    import oracle
    import torch
    rng = np.random.default_rng(11)
    d1 = (rng.random((5, 2048), dtype=np.float32) * 3e-4).astype(np.float32)
    d2 = (rng.random((5, 777), dtype=np.float32) * 3e-4).astype(np.float32)
    d1[2] += 1.0
    d2[2] += 1.0
    t1, t2 = torch.from_numpy(d1), torch.from_numpy(d2)
    p1 = torch.mean((t1 < 0.0001).float(), dim=1)
    p2 = torch.mean((t2 < 0.0001).float(), dim=1)
    f = 2 * p1 * p2 / (p1 + p2)
    f[torch.isnan(f)] = 0
    of, o1, o2 = oracle.fscore(d1, d2, mean="div")        # torch on the CPU: sum, then div_
    np.testing.assert_array_equal(o1, p1.numpy())
    np.testing.assert_array_equal(o2, p2.numpy())
    np.testing.assert_array_equal(of, f.numpy())
    assert of[2] == 0
    ff, f1, f2 = oracle.fscore(d1, d2)                     # the CUDA reduce kernel's arithmetic: within one ulp
    np.testing.assert_allclose(f2, o2, rtol=2e-7) and np.testing.assert_array_equal(f1, o1)   # 2048 is a power of two
