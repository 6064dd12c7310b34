#!/bin/bash
# Round-end evidence on one GPU box. Part "a": parity tests, the default bench line, the reference arm, step breakdowns,
# interleaved A/B of the round's switches. Part "b": ncu launch list of the bench command + full captures (each ncu pass only
# after the same command exited 0 without ncu); part "c": the attention-only subset of b.
# Usage: tools/gpu_final_round.sh <tag> a|b|c      logs: gpurun_out/<tag>_*
tag=${1:-final}; part=${2:-a}
mkdir -p gpurun_out
if [ "$part" == "a" ]; then
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/${tag}_smi.txt 2>&1
  timeout 600 python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/${tag}_tests.log 2>&1
  echo "tests rc=$?"; tail -n 2 gpurun_out/${tag}_tests.log
  cp gpurun_out/parity_report.json gpurun_out/${tag}_parity_report.json 2>/dev/null
  timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
  echo "bench rc=$?"; head -c 1500 gpurun_out/${tag}_bench_n1.json; echo
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
  echo "reference arm rc=$?"; cat gpurun_out/${tag}_bench_reference_arm.json
  timeout 300 python tools/profile_step.py 512 > gpurun_out/${tag}_step_breakdown.log 2>&1
  echo "profile_step rc=$?"; grep -v Warn gpurun_out/${tag}_step_breakdown.log | head -n 18 | cut -c1-150
  timeout 300 python tools/profile_step.py 64 > gpurun_out/${tag}_step_breakdown_b64.log 2>&1
  echo "profile_step 64 rc=$?"
  timeout 300 python tools/step_ab.py large 64 base nooverlap,nofusedbias,noqbias 2>&1 | grep -v Warn | tee gpurun_out/${tag}_step_ab_large_b64.log
elif [ "$part" == "c" ]; then
  # after a change to the attention kernels only: launch list + full captures of the attention forward / backward
  BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pairs 0 --sweep-images 0 --no-gpu-eager --skip-eager-roofline"
  $BCMD > gpurun_out/${tag}_ncu_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches.csv $BCMD > gpurun_out/${tag}_ncu_list.log 2>&1
  echo "ncu list rc=$?"
  NG="$BCMD --no-graph"
  $NG > gpurun_out/${tag}_ncu_plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_kd -s 13 -c 1 -o gpurun_out/${tag}_prof_attn_bwd $NG > gpurun_out/${tag}_ncu_full_attn.log 2>&1
  echo "ncu full (attention bwd) rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_persistent -s 13 -c 1 -o gpurun_out/${tag}_prof_attn_fwd $NG > gpurun_out/${tag}_ncu_full_attn_fwd.log 2>&1
  echo "ncu full (attention fwd) rc=$?"
  for r in prof_attn_bwd prof_attn_fwd; do
    python tools/summarize_ncu.py full gpurun_out/${tag}_$r.ncu-rep > gpurun_out/${tag}_ncu_full_$r.txt 2>&1
  done
  python tools/summarize_ncu.py launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_ncu_launch_list_step.txt 2>&1
  python tools/ncu_hot.py gpurun_out/${tag}_prof_attn_bwd.ncu-rep attention_bwd_kd 45 > gpurun_out/${tag}_ncu_attention_bwd_stall_sites.txt 2>&1
  rm -f gpurun_out/${tag}_prof_attn_fwd.ncu-rep gpurun_out/${tag}_prof_attn_bwd.ncu-rep
  ls -la gpurun_out/${tag}_*
else
  BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pairs 0 --sweep-images 0 --no-gpu-eager --skip-eager-roofline"
  $BCMD > gpurun_out/${tag}_ncu_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches.csv $BCMD > gpurun_out/${tag}_ncu_list.log 2>&1
  echo "ncu list rc=$?"
  NG="$BCMD --no-graph"
  $NG > gpurun_out/${tag}_ncu_plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 150 -c 6 -o gpurun_out/${tag}_prof_gemm $NG > gpurun_out/${tag}_ncu_full.log 2>&1
  echo "ncu full (gemm fwd) rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 250 -c 8 -o gpurun_out/${tag}_prof_gemm_bwd $NG > gpurun_out/${tag}_ncu_full_bwd.log 2>&1
  echo "ncu full (gemm bwd) rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_kd -s 13 -c 1 -o gpurun_out/${tag}_prof_attn_bwd $NG > gpurun_out/${tag}_ncu_full_attn.log 2>&1
  echo "ncu full (attention bwd) rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:layernorm_bwd -s 26 -c 2 -o gpurun_out/${tag}_prof_ln_bwd $NG > gpurun_out/${tag}_ncu_full_ln.log 2>&1
  echo "ncu full (layernorm bwd) rc=$?"
  for r in prof_gemm prof_gemm_bwd prof_attn_bwd prof_ln_bwd; do
    python tools/summarize_ncu.py full gpurun_out/${tag}_$r.ncu-rep > gpurun_out/${tag}_ncu_full_$r.txt 2>&1
  done
  python tools/summarize_ncu.py launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_ncu_launch_list_step.txt 2>&1
  python tools/ncu_hot.py gpurun_out/${tag}_prof_attn_bwd.ncu-rep attention_bwd_kd 45 > gpurun_out/${tag}_ncu_attention_bwd_stall_sites.txt 2>&1
  rm -f gpurun_out/${tag}_prof_gemm.ncu-rep gpurun_out/${tag}_prof_gemm_bwd.ncu-rep gpurun_out/${tag}_prof_ln_bwd.ncu-rep
  ls -la gpurun_out/${tag}_*
fi
