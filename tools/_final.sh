set -u
O=gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python - <<'PY'
import json, subprocess, sys
out = {"note": "tools/model_step.py, B = 32, one B200: zero_grad + forward + backward + Adam of the reference's unmodified models (staged under baseline/_ref/completion) on the reference's kernels (ref), on this library (ours), and with every opt-in patch of mvp_benchmark_b200.model_patches (ours_patched: kNN / top-k, fused sampling chains, loss epilogue, SA_module convolutions before the gather, neighbour aggregation, gather-max, thin 1x1 convolutions as matmuls); median step of 8 after 3 warm-up steps"}
def run(tag, args):
    p = subprocess.run([sys.executable, "tools/model_step.py"] + args, capture_output=True, text=True, timeout=900)
    line = [l for l in p.stdout.splitlines() if l.startswith("MODEL_STEP ")]
    return json.loads(line[-1][len("MODEL_STEP "):]) if line else {"error": p.stderr[-400:]}
for model, extra in (("vrcnet", []), ("ecg", []), ("pcn", ["--num-points", "16384"])):
    key = model if not extra else "pcn_num_points_16384"
    out[key] = {}
    for tag, a in (("ref", ["--ops", "ref"]), ("ours", ["--ops", "ours"]), ("ours_patched", ["--ops", "ours", "--patch-knn", "--profile", "--top", "30"])):
        if model == "pcn" and tag == "ours_patched":
            continue
        r = run(tag, ["--model", model, "--steps", "8", "--warmup", "3"] + a + extra)
        out[key][tag] = r
        print(key, tag, r.get("ms_per_step"), r.get("loss"), flush=True)
json.dump(out, open("gpurun_out/r2_model_step.json", "w"), indent=1)
PY
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; python -c "
import json; d=json.loads(open('$O/r2_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['extra'].get('vrcnet_training_step'))"
