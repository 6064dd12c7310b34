#!/bin/bash
# Runs the kernel-level GPU tests group by group, each in its own process and under `timeout`, so that a trapped
# kernel (sticky CUDA error) in one group cannot poison or hang the others. Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in "gemm_kmajor" "gemm_nobias or gemm_residual or gemm_gelu or gemm_strided or gemm_sumsq" "gemm_dgrad" "gemm_wgrad" "layernorm" "attention" "casts or im2col or assemble or colsum or rowsumsq"; do
  name=$(echo "$grp" | tr ' ' '_' | cut -c1-40)
  echo "=== $grp"
  timeout 420 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" -x --tb=short > "gpurun_out/k_${name}.log" 2>&1
  echo "rc=$?"
  tail -n 25 "gpurun_out/k_${name}.log"
done
