#!/bin/bash
# ncu evidence for one round (one GPU): launch list of the bench command + full captures of the dominant kernels.
# Each ncu pass runs only after the same command has exited 0 without ncu. Usage: tools/gpu_ncu_round.sh <tag>
tag=${1:-ncu}
mkdir -p gpurun_out
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
python tools/profile_sweep.py 64 large > gpurun_out/${tag}_sweep_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_perturb -s 3 -c 1 -o gpurun_out/${tag}_prof_perturb python tools/profile_sweep.py 64 large > gpurun_out/${tag}_ncu_full_perturb.log 2>&1
echo "ncu full (perturbation attention) rc=$?"
# summaries are made here, on the box: only they and the two small attention reports travel back (64 MiB limit)
for r in prof_gemm prof_gemm_bwd prof_attn_bwd prof_perturb; do
  python tools/summarize_ncu.py full gpurun_out/${tag}_$r.ncu-rep > gpurun_out/${tag}_ncu_full_$r.txt 2>&1
done
python tools/summarize_ncu.py launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_ncu_launch_list_step.txt 2>&1
python tools/ncu_hot.py gpurun_out/${tag}_prof_attn_bwd.ncu-rep attention_bwd_kd 45 > gpurun_out/${tag}_ncu_attention_bwd_stall_sites.txt 2>&1
python tools/ncu_hot.py gpurun_out/${tag}_prof_perturb.ncu-rep attention_perturb 45 > gpurun_out/${tag}_ncu_perturb_stall_sites.txt 2>&1
rm -f gpurun_out/${tag}_prof_gemm.ncu-rep gpurun_out/${tag}_prof_gemm_bwd.ncu-rep
ls -la gpurun_out/${tag}_*
