#!/bin/bash
# Attention-only GPU round: parity tests, timings, optional chunk timeline of the backward kernel.
tag=${1:-attn}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" -x --tb=short > gpurun_out/${tag}_tests.log 2>&1
echo "attention tests rc=$?"; tail -n 15 gpurun_out/${tag}_tests.log
timeout 300 python tools/attn_bench.py > gpurun_out/${tag}_attn_bench.log 2>&1
echo "attn_bench rc=$?"; tail -n 20 gpurun_out/${tag}_attn_bench.log
if [ "$2" == "dbg" ]; then
  VITB200_DBG_TIMING=1 timeout 120 python tools/attn_dbg.py > gpurun_out/${tag}_attn_dbg.log 2>&1; tail -n 40 gpurun_out/${tag}_attn_dbg.log
fi
