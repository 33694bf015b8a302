#!/bin/bash
# BASELINE configs C4 (VRCNet data-parallel step, B=32 per GPU) and C5 (EMD B=64 n=8192 per GPU) at N = 1, 2, 4, 8 on ONE box.
# Run under `gpurun --gpus 8`; writes gpurun_out/r2_model_scale.jsonl and gpurun_out/r2_emd_scale.jsonl.
set -u
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/r2_model_scale.jsonl
: > $OUT/r2_emd_scale.jsonl
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N))"; fi
  timeout 600 $L tools/model_step.py --model vrcnet --ops ours --steps 8 --warmup 3 2>$OUT/r2_model_scale_n$N.err | grep MODEL_STEP | sed 's/^MODEL_STEP //' >> $OUT/r2_model_scale.jsonl
  timeout 600 $L tools/model_step.py --model vrcnet --ops ours --patch-knn --steps 8 --warmup 3 2>>$OUT/r2_model_scale_n$N.err | grep MODEL_STEP | sed 's/^MODEL_STEP //' >> $OUT/r2_model_scale.jsonl
  timeout 300 $L tools/emd_scale.py 2>>$OUT/r2_model_scale_n$N.err | grep workload >> $OUT/r2_emd_scale.jsonl
done
nvidia-smi topo -m > $OUT/r2_topo.txt 2>&1
cat $OUT/r2_model_scale.jsonl | cut -c1-400
cat $OUT/r2_emd_scale.jsonl
