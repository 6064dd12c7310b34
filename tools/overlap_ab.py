"""A/B of the weight-gradient side stream (ops.wgrad_overlap) inside the graph-replayed training step.

Usage: python tools/overlap_ab.py [model] [batch ...]     e.g.  python tools/overlap_ab.py base 512 64
For every batch size: the step without the side stream, with it at lag 1 and lag 2; ms/step over 20 replays (CUDA events)
and the relative difference of the parameters after 5 steps from the same initialisation against the run without overlap
(split-K reduce-add order is not deterministic, so a few 1e-7 is what two identical runs differ by)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import build_model, ops  # noqa: E402
from vit_plasticity_b200.finetune import GraphedTrainStep, build_optimizer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "base"
batches = [int(a) for a in sys.argv[2:]] or [512, 64]
dev = "cuda"


def run(batch: int, overlap: bool, lag: int, steps: int = 20, prio: bool = True):
    os.environ["VB_WGRAD_STREAM"] = "1" if overlap else "0"
    os.environ["VB_GRAPH_PRIORITY"] = "1" if prio else "0"
    ops._Side.lag = lag
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
    l5, g5 = float(loss), float(gn)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        step([(x, y)])
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps, params, l5, g5


for b in batches:
    base_ms, base_p, bl, bg = run(b, False, 1)
    print(f"batch {b:4d}  no overlap       {base_ms:8.3f} ms/step   loss {bl:.6f} gnorm {bg:.6f}", flush=True)
    for lag, prio in ((1, True), (1, False), (2, True)):
        ms, p, l, gnorm = run(b, True, lag, prio=prio)
        rel = float((p - base_p).norm() / base_p.norm())
        print(f"batch {b:4d}  overlap lag {lag} prio {int(prio)}  {ms:8.3f} ms/step   loss {l:.6f} gnorm {gnorm:.6f}   params rel diff {rel:.2e}   ({base_ms / ms:.3f}x)", flush=True)
    ms2, p2, _, _ = run(b, False, 1)
    print(f"batch {b:4d}  no overlap again {ms2:8.3f} ms/step   params rel diff {float((p2 - base_p).norm() / base_p.norm()):.2e}", flush=True)
