timeout 600 python -m pytest tests/test_gpu_python_api.py -x -q 2>&1 | tail -3
for m in vrcnet ecg; do
timeout 600 python tools/model_step.py --model $m --ops ours --patch-knn --steps 8 --warmup 3 2>/dev/null | grep MODEL_STEP | sed 's/^MODEL_STEP //' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['model'], d['ms_per_step'], d['loss'])"
done
