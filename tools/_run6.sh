timeout 600 python -m pytest tests/test_gpu_python_api.py -x -q -k "sa_module or chains or weights or pointwise" 2>&1 | tail -4
timeout 600 python tools/model_step.py --model vrcnet --ops ours --patch-knn --steps 8 --warmup 3 --profile --top 30 2>gpurun_out/r2_ms.err | grep MODEL_STEP | sed 's/^MODEL_STEP //' > gpurun_out/r2_model_vrcnet_patched.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_model_vrcnet_patched.json'))
print(d['ms_per_step'], d['loss'])
for r in d['profile']['top'][:30]: print('  ', r['ms'], r['calls'], r['kernel'][:90])
PY
tail -3 gpurun_out/r2_ms.err
