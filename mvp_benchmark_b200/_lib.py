"""ctypes binding of libmvp_ops.so (include/mvp_ops.h) — the only bridge between the Python operator
layer and the sm_100a kernels.  There is NO CPU or PyTorch fallback: if the shared library is missing
the import fails, and every op refuses non-CUDA tensors.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# MVP_OPS_LIB: tuning / debugging aid (tools/pair_variants.py builds alternative libraries with other macro knobs)
LIB_PATH = os.environ.get("MVP_OPS_LIB") or os.path.join(_HERE, "libmvp_ops.so")

_c_int = ctypes.c_int
_c_float = ctypes.c_float
_c_size_t = ctypes.c_size_t
_p = ctypes.c_void_p


class MvpOpsError(RuntimeError):
    pass


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -m mvp_benchmark_b200.build` (needs nvcc; cross-compiles sm_100a without a GPU). "
            "There is no CPU fallback for these operators.")
    lib = ctypes.CDLL(LIB_PATH)
    sig = {
        "mvp_abi_version": (_c_int, []),
        "mvp_build_info": (ctypes.c_char_p, []),
        "mvp_error_string": (ctypes.c_char_p, [_c_int]),
        "mvp_launch_count": (ctypes.c_ulonglong, []),
        "mvp_chamfer_forward_workspace_bytes": (_c_size_t, [_c_int] * 3),
        "mvp_chamfer_forward": (_c_int, [_c_int] * 3 + [_p] * 7 + [_c_size_t, _p]),
        "mvp_chamfer_forward_algo": (_c_int, [_c_int] * 4 + [_p] * 7 + [_c_size_t, _p]),
        "mvp_chamfer_backward": (_c_int, [_c_int] * 3 + [_p] * 8 + [_p]),
        "mvp_chamfer_backward_algo": (_c_int, [_c_int] * 4 + [_p] * 8 + [_p]),
        "mvp_emd_forward_workspace_bytes": (_c_size_t, [_c_int] * 2),
        "mvp_emd_forward": (_c_int, [_c_int] * 3 + [_p, _p, _c_float, _c_int, _p, _p, _p, _c_size_t, _p]),
        "mvp_emd_forward_algo": (_c_int, [_c_int] * 4 + [_p, _p, _c_float, _c_int, _p, _p, _p, _c_size_t, _p]),
        "mvp_emd_backward": (_c_int, [_c_int] * 2 + [_p] * 5 + [_p]),
        "mvp_furthest_point_sampling": (_c_int, [_c_int] * 3 + [_p] * 3 + [_p]),
        "mvp_furthest_point_sampling_with_dist": (_c_int, [_c_int] * 3 + [_p] * 3 + [_p]),
        "mvp_ball_query": (_c_int, [_c_int] * 3 + [_c_float, _c_float, _c_int] + [_p] * 3 + [_p]),
        "mvp_gather_points": (_c_int, [_c_int] * 4 + [_p] * 3 + [_p]),
        "mvp_gather_points_grad": (_c_int, [_c_int] * 4 + [_p] * 3 + [_p]),
        "mvp_group_points": (_c_int, [_c_int] * 5 + [_p] * 3 + [_p]),
        "mvp_group_points_grad": (_c_int, [_c_int] * 5 + [_p] * 3 + [_p]),
        "mvp_three_nn": (_c_int, [_c_int] * 3 + [_p] * 4 + [_p]),
        "mvp_three_interpolate": (_c_int, [_c_int] * 4 + [_p] * 4 + [_p]),
        "mvp_three_interpolate_grad": (_c_int, [_c_int] * 4 + [_p] * 4 + [_p]),
        "mvp_scatter_workspace_bytes": (_c_size_t, [_c_int] * 3),
        "mvp_gather_points_grad_ws": (_c_int, [_c_int] * 4 + [_p] * 4 + [_c_size_t, _p]),
        "mvp_group_points_grad_ws": (_c_int, [_c_int] * 5 + [_p] * 4 + [_c_size_t, _p]),
        "mvp_three_interpolate_grad_ws": (_c_int, [_c_int] * 4 + [_p] * 5 + [_c_size_t, _p]),
        "mvp_three_nn_workspace_bytes": (_c_size_t, [_c_int] * 3),
        "mvp_three_nn_ws": (_c_int, [_c_int] * 3 + [_p] * 5 + [_c_size_t, _p]),
        "mvp_knn": (_c_int, [_c_int] * 4 + [_p] * 4 + [_p]),
        "mvp_three_nn_weights": (_c_int, [_c_int] * 2 + [_p] * 2 + [_p]),
        "mvp_gather_max": (_c_int, [_c_int] * 5 + [_p] * 4 + [_p]),
        "mvp_gather_max_grad": (_c_int, [_c_int] * 4 + [_p] * 3 + [_p]),
        "mvp_neighbor_weighted_sum": (_c_int, [_c_int] * 5 + [_p] * 4 + [_p]),
        "mvp_neighbor_weighted_sum_grad": (_c_int, [_c_int] * 5 + [_p] * 6 + [_p]),
        "mvp_pointwise_conv": (_c_int, [_c_int] * 4 + [_p] * 3 + [_c_int, _p] + [_p]),
        "mvp_pointwise_conv_masked": (_c_int, [_c_int] * 4 + [_p] * 4 + [_p]),
        "mvp_pointwise_wgrad_workspace_bytes": (_c_size_t, [_c_int] * 3),
        "mvp_pointwise_wgrad": (_c_int, [_c_int] * 4 + [_p] * 5 + [_c_size_t, _p]),
        "mvp_topk_rows_sqdist": (_c_int, [_c_int] * 3 + [_p] * 5 + [_p]),
        "mvp_max_last": (_c_int, [ctypes.c_longlong, _c_int] + [_p] * 3 + [_p]),
        "mvp_max_last_grad": (_c_int, [ctypes.c_longlong, _c_int] + [_p] * 3 + [_p]),
        "mvp_bias_add": (_c_int, [_c_int] * 3 + [_p] * 2 + [_c_int, _p]),
        "mvp_channel_sum_workspace_bytes": (_c_size_t, [_c_int] * 2),
        "mvp_channel_sum": (_c_int, [_c_int] * 3 + [_p] * 3 + [_c_size_t, _p]),
        "mvp_topk_rows": (_c_int, [ctypes.c_longlong, _c_int, _c_int] + [_p] * 4 + [_p]),
        "mvp_three_nn_weights_ws": (_c_int, [_c_int] * 3 + [_p] * 6 + [_c_size_t, _p]),
        "mvp_furthest_point_sampling_gather": (_c_int, [_c_int] * 3 + [_p] * 4 + [_c_int, _p]),
        "mvp_ball_query_group": (_c_int, [_c_int] * 3 + [_c_float] * 2 + [_c_int] + [_p] * 4 + [_p]),
        "mvp_fscore": (_c_int, [_c_int] * 3 + [_p] * 2 + [_c_float] + [_p] * 3 + [_p]),
        "mvp_chamfer_loss": (_c_int, [_c_int] * 3 + [_p] * 4 + [_p]),
        "mvp_chamfer_loss_grad": (_c_int, [_c_int] * 3 + [_p] * 6 + [_p]),
        "mvp_knn_points_workspace_bytes": (_c_size_t, [_c_int] * 4),
        "mvp_knn_points": (_c_int, [_c_int] * 4 + [_p] * 5 + [_c_size_t, _p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib, tuple(sig)


lib, EXPORTS = _load()


def check(rc, what):
    if rc != 0:
        raise MvpOpsError(f"{what} failed with code {rc}: {lib.mvp_error_string(rc).decode()}")


def launch_count():
    return int(lib.mvp_launch_count())


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream_of(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda(*tensors, dtype=None, what="mvp op"):
    for t in tensors:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{what}: expected a torch.Tensor, got {type(t).__name__}")
        if not t.is_cuda:
            raise MvpOpsError(f"{what}: CUDA tensors only — these operators have no CPU path "
                              f"(got a tensor on {t.device})")
        if dtype is not None and t.dtype != dtype:
            raise TypeError(f"{what}: expected dtype {dtype}, got {t.dtype}")
    dev = tensors[0].device
    for t in tensors[1:]:
        if t.device != dev:
            raise MvpOpsError(f"{what}: all tensors must be on the same device ({dev} vs {t.device})")
    return dev


def workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)
