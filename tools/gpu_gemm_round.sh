#!/bin/bash
# GEMM-only GPU round: microbenchmarks, the GEMM parity tests under both tile mappings, GEMM timings under both.
tag=${1:-gemm}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
if [ -x tools/microbench/gelu_rate ] && [ "$2" == "micro" ]; then timeout 120 tools/microbench/gelu_rate > gpurun_out/${tag}_gelu_rate.log 2>&1; echo "gelu_rate rc=$?"; cat gpurun_out/${tag}_gelu_rate.log; fi
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" -x --tb=short > gpurun_out/${tag}_tests.log 2>&1
echo "gemm tests rc=$?"; tail -n 30 gpurun_out/${tag}_tests.log
KB_ONLY=gemm KB_GEMM_MODES=${KB_GEMM_MODES:-1230} timeout 900 python tools/kernel_bench.py > gpurun_out/${tag}_kb_modes.log 2>&1
echo "kernel_bench modes rc=$?"; cat gpurun_out/${tag}_kb_modes.log
