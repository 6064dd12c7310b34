#!/bin/bash
# One GPU-box round: parity tests -> bench -> (optional) ncu launch list + ncu --set full of the dominant kernel.
# Usage: tools/gpu_round.sh <tag> [ncu]        logs: gpurun_out/<tag>_*.log
tag=${1:-round}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/${tag}_tests.log
timeout 600 python tools/kernel_bench.py > gpurun_out/${tag}_kernel_bench.log 2>&1
echo "kernel_bench rc=$?"
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
timeout 600 python tools/profile_step.py > gpurun_out/${tag}_step_breakdown.log 2>&1
echo "profile_step rc=$?"; head -n 16 gpurun_out/${tag}_step_breakdown.log
if [ "$2" == "ncu" ]; then
  BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pairs 0"
  $BCMD > gpurun_out/${tag}_ncu_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches.csv $BCMD > gpurun_out/${tag}_ncu_list.log 2>&1
  echo "ncu list rc=$?"
  $BCMD > gpurun_out/${tag}_ncu_plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 150 -c 6 -o gpurun_out/${tag}_prof_gemm $BCMD > gpurun_out/${tag}_ncu_full.log 2>&1
  echo "ncu full rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_ -s 13 -c 2 -o gpurun_out/${tag}_prof_attn_bwd $BCMD > gpurun_out/${tag}_ncu_full_attn.log 2>&1
  echo "ncu full (attention bwd) rc=$?"
fi
# per-kernel regression guard: this run's kernel_bench log against the last kept profiles/*kernel_bench*.log, both normalised
# by their torch controls (run it here or on the CPU box; exit code 1 = some kernel is > 5 % slower)
python tools/kernel_regression.py gpurun_out/${tag}_kernel_bench.log; echo "kernel regression guard rc=$?"
