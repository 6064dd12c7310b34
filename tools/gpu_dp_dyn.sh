#!/bin/bash
# 8-GPU A/B: static tile stride + 256 MB buckets (default) against work-stealing GEMM tiles + 64 MB buckets
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --pairs 0 > gpurun_out/r2j_${name}.json 2> gpurun_out/r2j_${name}.err
  echo "$name rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/r2j_${name}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['us_per_launch'])" 2>&1 | tail -1)"; }
run steal_64mb VB_GEMM_DYNAMIC=1 VB_DP_BUCKET_MB=64
run static_256mb VB_GEMM_DYNAMIC=0
