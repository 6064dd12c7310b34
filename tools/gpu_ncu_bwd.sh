#!/bin/bash
# ncu --set full of the BACKWARD GEMM launches of one block (wgrad fc2, fc2 dgrad x gelu', wgrad fc1, dgrad fc1, wgrad proj,
# proj dgrad + delta, wgrad qkv, dgrad qkv) and of one attention backward launch. Usage: tools/gpu_ncu_bwd.sh <tag>
tag=${1:-bwd}
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pairs 0"
$BCMD > gpurun_out/${tag}_plain.log 2>&1 || exit 1
# 3 warm-up steps x 146 GEMM launches + the 49 forward launches of the timed step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 487 -c 8 -o gpurun_out/${tag}_prof_gemm_bwd $BCMD > gpurun_out/${tag}_ncu_gemm_bwd.log 2>&1
echo "ncu gemm bwd rc=$?"
$BCMD > gpurun_out/${tag}_plain2.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_ -s 37 -c 1 -o gpurun_out/${tag}_prof_attn_bwd $BCMD > gpurun_out/${tag}_ncu_attn_bwd.log 2>&1
echo "ncu attn bwd rc=$?"
