timeout 600 python tools/model_step.py --model vrcnet --ops ours --patch-knn --steps 8 --warmup 3 --profile --top 40 2>/dev/null | grep MODEL_STEP | sed 's/^MODEL_STEP //' > gpurun_out/r2_model_vrcnet_patched.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_model_vrcnet_patched.json'))
print(d['ms_per_step'], d['wall_ms_per_step'], d['profile']['cuda_ms_total'])
for r in d['profile']['top'][:40]: print('  ', r['ms'], r['calls'], r['kernel'][:110])
PY
