#!/bin/bash
# Bucket-size A/B of the data-parallel all-reduce at the strong-scaling split (64 images per GPU): tools/gpu_n8_buckets.sh <tag> [N]
tag=${1:-bk}; N=${2:-8}; port=29600
for cfg in "VB_DP_BUCKET_MB=32" "VB_DP_BUCKET_MB=64" "VB_DP_BUCKET_MB=128" "VB_DP_BUCKET_MB=256" "VB_DP_OVERLAP=0"; do
  port=$((port+1))
  env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 3 --pairs 0 --sweep-images 0 --no-e2e --skip-eager-roofline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', d['value'], d['ms_per_step'])" | tee -a gpurun_out/${tag}_buckets.log
done
