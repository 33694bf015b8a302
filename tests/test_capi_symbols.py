"""CPU-only: libmvp_ops.so loads and exports exactly what include/mvp_ops.h declares."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mvp_ops.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    names = _declared()
    for need in ["mvp_chamfer_forward", "mvp_chamfer_backward", "mvp_emd_forward", "mvp_emd_backward",
                 "mvp_furthest_point_sampling", "mvp_furthest_point_sampling_with_dist", "mvp_ball_query",
                 "mvp_gather_points", "mvp_gather_points_grad", "mvp_group_points", "mvp_group_points_grad",
                 "mvp_three_nn", "mvp_three_interpolate", "mvp_three_interpolate_grad", "mvp_knn"]:
        assert need in names


def test_library_exports_every_declared_symbol():
    from mvp_benchmark_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in mvp_ops.h but not exported"
    assert sorted(_lib.EXPORTS) == _declared()          # the ctypes table covers the header too


def test_library_is_sm100a_only_and_has_no_torch_dependency():
    from mvp_benchmark_b200 import _lib
    archs, err = set(), ""
    for _ in range(3):  # cuobjdump now and then returns nothing on a loaded machine: ask again before judging
        r = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
        archs, err = set(re.findall(r"sm_\d+a?", r.stdout)), r.stderr
        if archs:
            break
    assert archs == {"sm_100a"}, (archs, err[-500:])
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "python" not in ldd and "c10" not in ldd


def test_identification_and_error_text():
    from mvp_benchmark_b200 import _lib
    assert _lib.lib.mvp_abi_version() == 1
    assert b"sm_100a" in _lib.lib.mvp_build_info()
    assert b"multiple of 1024" in _lib.lib.mvp_error_string(-4)
    assert _lib.lib.mvp_error_string(0) == b"ok"


def test_argument_validation_needs_no_gpu():
    # every entry point validates sizes before touching the device
    from mvp_benchmark_b200 import _lib
    L = _lib.lib
    z = ctypes.c_void_p(0)
    assert L.mvp_emd_forward(1, 1000, 1000, z, z, 0.005, 1, z, z, z, 0, z) == -4
    assert L.mvp_emd_forward(1, 1024, 2048, z, z, 0.005, 1, z, z, z, 0, z) == -2
    assert L.mvp_emd_forward(513, 1024, 1024, z, z, 0.005, 1, z, z, z, 0, z) == -3
    assert L.mvp_knn(1, 8, 8, 101, z, z, z, z, z) == -1
    assert L.mvp_chamfer_forward(-1, 8, 8, z, z, z, z, z, z, z, 0, z) == -1
    assert L.mvp_chamfer_forward(0, 8, 8, z, z, z, z, z, z, z, 0, z) == 0     # empty batch is a no-op
    assert L.mvp_gather_points(2, 0, 8, 8, z, z, z, z) == 0
    # the 1x1-layer entries (csrc/pointwise.cu) and the neighbour max
    assert L.mvp_pointwise_conv(1, 0, 8, 8, z, z, z, 0, z, z) == -1
    assert L.mvp_pointwise_conv(1, 8, 8, 8, z, z, z, 0, z, z) == -1            # null operands
    assert L.mvp_pointwise_conv(0, 8, 8, 8, z, z, z, 0, z, z) == 0             # empty batch
    assert L.mvp_pointwise_conv_masked(1, 8, 8, 8, z, z, z, z, z) == -1
    assert L.mvp_pointwise_wgrad(1, 256, 8, 8, z, z, z, ctypes.c_void_p(16), z, 0, z) == -1   # 256 channels + the ones row
    assert L.mvp_pointwise_wgrad_workspace_bytes(64, 256, 1) == 2 * 74 * 128 * 80 * 4
    assert L.mvp_pointwise_wgrad_workspace_bytes(300, 8, 0) == 0
    assert L.mvp_bias_add(1, 0, 8, z, z, 0, z) == -1 and L.mvp_bias_add(0, 4, 8, z, z, 0, z) == 0
    assert L.mvp_channel_sum(1, 4, 8, ctypes.c_void_p(16), ctypes.c_void_p(16), z, 0, z) == -5    # workspace missing
    assert L.mvp_channel_sum_workspace_bytes(64, 1024) == 1024 * 4 and L.mvp_channel_sum_workspace_bytes(64, 4) == 4 * 64 * 4
    assert L.mvp_max_last(4, 0, z, z, z, z) == -1 and L.mvp_max_last(4, 256, z, z, z, z) == -1 and L.mvp_max_last(0, 16, z, z, z, z) == 0
    assert L.mvp_fscore(1, 0, 8, z, z, 1e-4, z, z, z, z) == -1 and L.mvp_fscore(0, 8, 8, z, z, 1e-4, z, z, z, z) == 0
    assert L.mvp_topk_rows_sqdist(1, 8, 9, z, z, z, z, z, z) == -1 and L.mvp_topk_rows_sqdist(0, 8, 4, z, z, z, z, z, z) == 0
