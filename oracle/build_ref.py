#!/usr/bin/env python
"""TEST INFRASTRUCTURE — build recipe for oracle/_ref/libref_ops.so.

Compiles the REFERENCE's nine hot-path .cu files, unmodified and from where they lie under
/root/reference, for sm_100a, together with oracle/ref_capi.cu (our thin C-ABI shim), into
oracle/_ref/libref_ops.so.  Outputs go only into oracle/_ref/ (git-ignored, NOT gpurun-ignored, so
the built library travels to the GPU box).  Never copies reference sources.

The reference's own build system is not used (utils/mm3d_pn2/setup.py needs mmcv and THC).
Runs only where /root/reference exists (this container); on the GPU box the prebuilt .so is used.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MVP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

REF_SOURCES = [
    "utils/metrics/CD/chamfer3D/chamfer3D.cu",
    "utils/metrics/EMD/emd_cuda.cu",
    "utils/mm3d_pn2/ops/furthest_point_sample/src/furthest_point_sample_cuda.cu",
    "utils/mm3d_pn2/ops/ball_query/src/ball_query_cuda.cu",
    "utils/mm3d_pn2/ops/gather_points/src/gather_points_cuda.cu",
    "utils/mm3d_pn2/ops/group_points/src/group_points_cuda.cu",
    "utils/mm3d_pn2/ops/interpolate/src/three_nn_cuda.cu",
    "utils/mm3d_pn2/ops/interpolate/src/three_interpolate_cuda.cu",
    "utils/mm3d_pn2/ops/knn/src/knn_cuda.cu",
]


def available():
    return os.path.isdir(REF) and all(os.path.isfile(os.path.join(REF, s)) for s in REF_SOURCES)


def lib_path():
    return os.path.join(OUT, "libref_ops.so")


def build(force=False, verbose=False):
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(REF, s) for s in REF_SOURCES] + [os.path.join(HERE, "ref_capi.cu")]
    target = lib_path()
    if not force and os.path.isfile(target) and all(
            os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs + [__file__]):
        return target
    import torch
    from torch.utils import cpp_extension as ce
    inc = []
    for p in ce.include_paths():
        inc += ["-isystem", p]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    # Same flags the reference's JIT `load` would give (no fast-math, default -fmad=true), plus the
    # arch the reference never names (utils/mm3d_pn2/setup.py:43-47 passes no -gencode).
    common = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
              "-w"] + inc
    objs = []

    def cc(src):
        obj = os.path.join(OUT, os.path.basename(src).replace(".cu", ".o"))
        cmd = common + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(cc, srcs))
    link = ["nvcc", "-shared", "-o", target] + objs + [
        "-L" + torch_lib, "-lc10", "-ltorch_cpu", "-ltorch_cuda", "-lc10_cuda", "-ltorch",
        "-Xlinker", "-rpath," + torch_lib]
    if verbose:
        print(" ".join(link), flush=True)
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    return target


MODEL_FILES = [
    "completion/model_utils.py", "completion/models/__init__.py", "completion/models/pcn.py",
    "completion/models/ecg.py", "completion/models/vrcnet.py", "completion/cfgs/pcn.yaml",
    "completion/cfgs/ecg.yaml", "completion/cfgs/vrcnet.yaml",
]


def stage_models():
    """Stage the reference's completion MODELS (the callers of the hot path: SURVEY.md §2.1 row 9, out of
    scope and not rebuilt) unmodified under baseline/_ref/completion/ — the git-ignored place for an
    unmodified copy of the reference, which travels to the GPU box but never enters the history — for
    tools/model_step.py, which times an unmodified VRCNet / PCN / ECG training step on our operators and
    on the reference kernels on the same B200 (/root/reference does not exist on the GPU box).
    oracle/_ref/ itself holds build outputs only."""
    import shutil
    dest = os.path.join(os.path.dirname(HERE), "baseline", "_ref")
    for f in MODEL_FILES:
        dst = os.path.join(dest, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), dst)
    return os.path.join(dest, "completion")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(stage_models())
