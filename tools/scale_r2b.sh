#!/bin/bash
# Lean repeat of tools/scale_r2.sh after the reducer and patch changes: VRCNet step at N = 1 and N = 8 on one box.
set -u
O=gpurun_out
: > $O/r2_model_scale_b.jsonl
for N in 1 8; do
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N))"; fi
  for extra in "" "--patch-knn"; do
    timeout 600 $L tools/model_step.py --model vrcnet --ops ours $extra --steps 8 --warmup 3 2>>$O/r2_model_scale_b.err | grep MODEL_STEP | sed 's/^MODEL_STEP //' >> $O/r2_model_scale_b.jsonl
  done
done
cut -c1-330 $O/r2_model_scale_b.jsonl
