for args in "--patch-knn" ""; do
for N in 1 2; do
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
  timeout 600 $L tools/model_step.py --model vrcnet --ops ours $args --steps 8 --warmup 3 2>/dev/null | grep MODEL_STEP | sed 's/^MODEL_STEP //' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['patch_knn'], round(d['ms_per_step'],2), d['loss'], d.get('allreduce',{}).get('standalone_ms'))"
done; done
