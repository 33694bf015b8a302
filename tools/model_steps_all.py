#!/usr/bin/env python
"""Regenerate profiles/rN_model_step.json on a GPU box: for VRCNet, ECG and PCN (16 384 output points) the training step of
the reference's unmodified models on the reference's kernels (ref), on this library (ours) and with every opt-in patch of
mvp_benchmark_b200.model_patches (ours_patched, with its per-kernel profile).
    python tools/model_steps_all.py gpurun_out/model_step.json"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
NOTE = ("tools/model_step.py, B = 32, one B200: zero_grad + forward + backward + Adam of the reference's unmodified models "
        "(staged under baseline/_ref/completion) on the reference's kernels (ref), on this library (ours), and with every opt-in "
        "patch of mvp_benchmark_b200.model_patches (ours_patched: kNN / top-k without the score matrix, fused sampling chains, loss "
        "epilogue, SA_module convolutions before the gather, neighbour aggregation, gather-max, neighbour max, get_graph_feature, "
        "1x1 layers on the tcgen05 kernels / library GEMM + our bias kernels, PCN decoder's per-cloud share of conv1); median step "
        "of 8 after 3 warm-up steps")


def arm(model, extra):
    cmd = [sys.executable, os.path.join(HERE, "model_step.py"), "--model", model, "--steps", "8", "--warmup", "3", "--top", "40"] + extra
    p = subprocess.run(cmd, capture_output=True, text=True)
    line = [l for l in p.stdout.splitlines() if l.startswith("MODEL_STEP ")]
    if p.returncode != 0 or not line:
        return {"error": (p.stderr or p.stdout)[-1500:]}
    return json.loads(line[-1][len("MODEL_STEP "):])


def main():
    out = {"note": NOTE}
    for key, model, extra in (("vrcnet", "vrcnet", []), ("ecg", "ecg", []), ("pcn_num_points_16384", "pcn", ["--num-points", "16384"])):
        out[key] = {"ref": arm(model, extra + ["--ops", "ref"]), "ours": arm(model, extra + ["--ops", "ours"]),
                    "ours_patched": arm(model, extra + ["--ops", "ours", "--patch-knn", "--profile"])}
        r = out[key]
        if all("ms_per_step" in r[a] for a in r):
            r["speedup_vs_reference_kernels"] = {"ours": r["ref"]["ms_per_step"] / r["ours"]["ms_per_step"],
                                                 "ours_patched": r["ref"]["ms_per_step"] / r["ours_patched"]["ms_per_step"]}
        print(key, {a: r[a].get("ms_per_step", r[a]) for a in ("ref", "ours", "ours_patched")}, flush=True)
    with open(sys.argv[1] if len(sys.argv) > 1 else "model_step.json", "w") as f:
        f.write(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
