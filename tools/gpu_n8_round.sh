#!/bin/bash
# One 8-GPU box round (BASELINE.json configs[2], [3], [4] at N = 8): usage tools/gpu_n8_round.sh <tag> [N]
tag=${1:-n8}
N=${2:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29521 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_base_strong.json 2> gpurun_out/${tag}_bench_base_strong.err; echo "base strong rc=$?"
run 29522 --workload sweep --images 65536 > gpurun_out/${tag}_sweep_65536.json 2> gpurun_out/${tag}_sweep_65536.err; echo "sweep rc=$?"
run 29523 --model large --steps 10 --warmup 3 --pairs 0 --sweep-images 0 --no-e2e > gpurun_out/${tag}_bench_large_strong.json 2> gpurun_out/${tag}_bench_large_strong.err; echo "large strong rc=$?"
run 29524 --components emb,attn_norm,ffn_norm,ffn_fc1,ffn_fc2 --steps 10 --warmup 3 --pairs 0 --sweep-images 0 --no-e2e > gpurun_out/${tag}_bench_attention_only.json 2> gpurun_out/${tag}_bench_attention_only.err; echo "attention-only rc=$?"
run 29525 --components emb,attn_norm,mha,ffn_norm --steps 10 --warmup 3 --pairs 0 --sweep-images 0 --no-e2e > gpurun_out/${tag}_bench_mlp_only.json 2> gpurun_out/${tag}_bench_mlp_only.err; echo "mlp-only rc=$?"
for f in gpurun_out/${tag}_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("unparsable", e); sys.exit(0)
keep = {k: d.get(k) for k in ("metric", "value", "ms_per_step", "n_gpus", "scaling", "seconds", "gather_seconds", "pairs")}
keep["config"] = d.get("config", {}).get("workload", "")[:90]
keep["roofline"] = {k: d.get("roofline", {}).get(k) for k in ("frac", "whole_step_frac", "gemm_share_of_step")}
keep["dp_parity"] = d.get("dp_parity")
keep["e2e"] = d.get("e2e")
print(json.dumps(keep))
PY
done
tail -n 3 gpurun_out/${tag}_*.err | grep -v "OMP_NUM\|\*\*\*\*" | tail -20
