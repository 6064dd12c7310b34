#!/bin/bash
# Data-parallel knobs, back to back on the same N-GPU box: overlapped buckets (default) / one all-reduce after backward /
# fewer NCCL CTAs / other bucket sizes. Usage: tools/gpu_dp_sweep.sh <tag> <N>
tag=${1:-dp}; N=${2:-8}
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --pairs 0 > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.loads(open('gpurun_out/${tag}_${name}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])" 2>&1 | tail -1)"
}
run overlap64 VB_DP_OVERLAP=1
run nooverlap VB_DP_OVERLAP=0
run overlap64_cta8 VB_DP_OVERLAP=1 NCCL_MAX_CTAS=8
run overlap256 VB_DP_OVERLAP=1 VB_DP_BUCKET_MB=256
run nooverlap_cta16 VB_DP_OVERLAP=0 NCCL_MAX_CTAS=16
