"""Kernel-time breakdown of one finetuning step (torch.profiler / CUPTI); prints the top kernels by total time.
The weight-gradient side stream is switched off here (VB_WGRAD_STREAM=0 unless the variable is already set): kernels that run
concurrently share the SMs, so their individual durations would no longer add up to the step."""
import os
import sys

os.environ.setdefault("VB_WGRAD_STREAM", "0")
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import build_model  # noqa: E402
from vit_plasticity_b200.finetune import build_optimizer, train_step  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
model = build_model({"implementation": "vit", "model_name": "base", "pretrained": False, "in21k": True, "finetuning": True, "n_classes": 10}, device="cuda")
model.train()
opt = build_optimizer(model, "sgd", lr=1e-2, momentum=0.9, fused=True)
x = torch.randn(B, 3, 224, 224, device="cuda")
y = torch.randint(0, 10, (B,), device="cuda")
for _ in range(3):
    train_step(model, opt, [(x, y)], grad_clip=1.0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        train_step(model, opt, [(x, y)], grad_clip=1.0)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 2e3, e.count // 2) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device kernel time per step: {tot:.2f} ms")
for k, ms, n in rows[:40]:
    print(f"{ms:9.3f} ms {100*ms/tot:5.1f}%  x{n:<5d} {k[:110]}")
