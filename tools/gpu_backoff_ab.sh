#!/bin/bash
# A/B on one box: mbarrier wait loops with / without a nanosleep back-off in the attention kernels (rebuilt on the box).
mkdir -p gpurun_out
echo "== default (no back-off)"; python tools/attn_bench.py
for ns in 20 100; do
  rm -f vit_plasticity_b200/csrc/attention_tc3.o
  make -C vit_plasticity_b200/csrc EXTRA=-DVB_MBAR_BACKOFF_NS=$ns > gpurun_out/backoff_build_$ns.log 2>&1 || { echo build failed; tail -5 gpurun_out/backoff_build_$ns.log; }
  echo "== back-off $ns ns"; python tools/attn_bench.py
done
