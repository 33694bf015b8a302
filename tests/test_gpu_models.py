"""GPU: the reference's UNMODIFIED completion models (the callers of the hot path) take one
training step on top of this repository's operators, and the loss of that step equals the loss of the same step on
the reference's own CUDA kernels (oracle/ref_packages -> oracle/_ref/libref_ops.so), same seed, same weights, same
inputs.  The models are not part of the repository: oracle/build_ref.py stages them under the git-ignored
baseline/_ref/completion/ where /root/reference exists; without them (or without the reference library) the tests skip.
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref", "completion", "models", "vrcnet.py")
REFLIB = os.path.join(ROOT, "oracle", "_ref", "libref_ops.so")


def _step(model, *extra):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "model_step.py"), "--model", model, "--batch", "4", "--steps", "1",
           "--warmup", "0", *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in p.stdout.splitlines() if l.startswith("MODEL_STEP ")]
    assert p.returncode == 0 and lines, (p.stdout + p.stderr)[-3000:]
    return json.loads(lines[-1][len("MODEL_STEP "):])


@pytest.mark.skipif(not (os.path.isfile(STAGED) and os.path.isfile(REFLIB)), reason="reference models / kernels not staged")
@pytest.mark.parametrize("model", ["vrcnet", "ecg", "pcn"])
def test_first_training_step_matches_the_reference_kernels(cuda, model):
    ours, ref = _step(model, "--ops", "ours"), _step(model, "--ops", "ref")
    assert ours["params"] == ref["params"] > 1_000_000
    assert ours["loss"] == pytest.approx(ref["loss"], rel=2e-5), (ours["loss"], ref["loss"])


@pytest.mark.skipif(not os.path.isfile(STAGED), reason="reference models not staged")
@pytest.mark.parametrize("model", ["vrcnet", "ecg", "pcn"])
def test_step_with_the_opt_in_patches(cuda, model):
    """Every opt-in patch of model_patches applied (tools/model_step.py --patch-knn): the step runs and the loss of the
    first training step stays that of the unpatched model to 1e-3 (near-tie differences of the neighbour ranking, fp32
    instead of TF32 in the thin convolutions)."""
    plain, patched = _step(model, "--ops", "ours"), _step(model, "--ops", "ours", "--patch-knn")
    assert patched["patch_knn"] is True
    assert patched["loss"] == pytest.approx(plain["loss"], rel=1e-3), (patched["loss"], plain["loss"])
