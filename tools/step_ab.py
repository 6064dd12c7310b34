"""Interleaved A/B of training-step variants (graph-replayed, one GPU): every variant is built and captured first, then
timed round-robin (R rounds of S replays each), so box state / clocks / thermals hit all variants alike.

Usage: python tools/step_ab.py <model> <batch> <variant> [<variant> ...]
A variant is a comma-separated list of switches: base (nothing), nooverlap (weight gradients on the main stream), noprio
(graph captured on a default-priority stream), tile256 (no 192-column GEMM tiles), nofusedbias (stand-alone column sums
for the proj / fc2 bias gradients), noqbias (all three thirds of the qkv bias gradient inside the attention backward), steal
(work-stealing GEMM tile scheduler), lag<N> (blocks the side stream may trail the main chain by).
Prints ms/step per variant and round, the median, and the parameter difference after 5 steps against the first variant."""
import os
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import _lib as L  # noqa: E402
from vit_plasticity_b200 import build_model, ops  # noqa: E402
from vit_plasticity_b200.finetune import GraphedTrainStep, build_optimizer  # noqa: E402

name, batch = sys.argv[1], int(sys.argv[2])
variants = sys.argv[3:] or ["base", "nooverlap"]
ROUNDS, STEPS = int(os.environ.get("AB_ROUNDS", "4")), int(os.environ.get("AB_STEPS", "10"))
dev = "cuda"


def build(variant: str):
    sw = set(variant.split(","))
    os.environ["VB_WGRAD_STREAM"] = "0" if "nooverlap" in sw else "1"
    os.environ["VB_GRAPH_PRIORITY"] = "0" if "noprio" in sw else "1"
    ops._Side.lag = next((int(t[3:]) for t in sw if t.startswith("lag")), ops._Side.lag)
    ops._FUSED_BIAS = "nofusedbias" not in sw
    ops._QBIAS = "noqbias" not in sw
    L.lib().vb_set_gemm_tile_n(256 if "tile256" in sw else 0)
    L.lib().vb_set_gemm_scheduler(1 if "steal" in sw else 0)
    torch.manual_seed(0)
    model = build_model({"implementation": "vit", "model_name": name, "pretrained": False, "in21k": True, "finetuning": True, "n_classes": 10}, device=dev)
    model.train()
    opt = build_optimizer(model, "sgd", lr=1e-2, momentum=0.9, fused=True)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(batch, 3, 224, 224, device=dev, generator=g)
    y = torch.randint(0, 10, (batch,), device=dev, generator=g)
    step = GraphedTrainStep(model, opt, 1.0)
    for _ in range(5):
        loss, gn = step([(x, y)])
    torch.cuda.synchronize()
    params = torch.cat([p.detach().flatten() for p in model.parameters()]).clone()
    return {"name": variant, "step": step, "xy": (x, y), "params": params, "loss": float(loss), "gn": float(gn), "ms": [], "launches": step.launches_per_step}


def time_steps(v) -> float:
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(STEPS):
        v["step"]([v["xy"]])
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / STEPS


vs = [build(v) for v in variants]
for r in range(ROUNDS):
    for v in vs:
        v["ms"].append(time_steps(v))
print(f"model {name} batch {batch}: {ROUNDS} rounds x {STEPS} graph replays, round-robin")
for v in vs:
    rel = float((v["params"] - vs[0]["params"]).norm() / vs[0]["params"].norm())
    med = statistics.median(v["ms"])
    print(f"{v['name']:28s} median {med:8.3f} ms/step  ({statistics.median(vs[0]['ms']) / med:.3f}x)  rounds {' '.join(f'{m:.3f}' for m in v['ms'])}"
          f"  launches/step {v['launches']}  loss@5 {v['loss']:.5f} gnorm@5 {v['gn']:.5f}  params rel diff vs first {rel:.1e}", flush=True)
