#!/bin/bash
# BASELINE.json configs other than the headline one, single-GPU shapes; plus the reference arm smoke.
tag=${1:-cfg}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err; echo "$name rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/${tag}_${name}.json').read().strip().splitlines()[-1]); print(d.get('value'), d.get('unit'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('plasticity') or {}).get('value'))" 2>&1 | tail -1)"; }
run large_b64 --model large --batch 64 --steps 10 --warmup 3 --no-cpu-baseline --pairs 32
run attn_only --components emb,attn_norm,ffn_norm,ffn_fc1,ffn_fc2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --pairs 0
run mlp_only --components emb,attn_norm,mha,ffn_norm --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --pairs 0
run reference_arm --impl reference --steps 2 --warmup 1
