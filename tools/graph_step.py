"""Experiment: eager train_step vs the same step captured in one CUDA graph (whole-network capture)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import build_model  # noqa: E402
from vit_plasticity_b200.finetune import build_optimizer, train_step  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
name = sys.argv[2] if len(sys.argv) > 2 else "base"
dev = "cuda"
model = build_model({"implementation": "vit", "model_name": name, "pretrained": False, "in21k": True, "finetuning": True, "n_classes": 10}, device=dev)
model.train()
opt = build_optimizer(model, "sgd", lr=1e-2, momentum=0.9, fused=True)
x = torch.randn(B, 3, 224, 224, device=dev)
y = torch.randint(0, 10, (B,), device=dev)


def timed(fn, n=10):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for _ in range(3):
    train_step(model, opt, [(x, y)], grad_clip=1.0)
print(f"eager: {timed(lambda: train_step(model, opt, [(x, y)], grad_clip=1.0)):.2f} ms/step")

# whole-step capture: warm up on a side stream, then capture fwd + bwd + clip + optimizer into one graph
sx, sy = x.clone(), y.clone()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        train_step(model, opt, [(sx, sy)], grad_clip=1.0)
torch.cuda.current_stream().wait_stream(side)
opt.zero_grad(set_to_none=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss, gnorm = train_step(model, opt, [(sx, sy)], grad_clip=1.0)
g.replay()
torch.cuda.synchronize()
print("graph loss", float(loss), "gnorm", float(gnorm))
print(f"graph: {timed(g.replay):.2f} ms/step")
l0 = float(loss)
for _ in range(20):
    g.replay()
torch.cuda.synchronize()
print("loss after 20 more graph steps", float(loss), "(was", l0, ")")
