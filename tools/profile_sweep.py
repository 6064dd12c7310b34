"""Kernel-time breakdown of one sweep chunk (64 images x the eps grid) of the fused estimator (CUPTI via torch.profiler)."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import build_model  # noqa: E402
from vit_plasticity_b200.plasticity import PlasticityEstimator  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
name = sys.argv[2] if len(sys.argv) > 2 else "large"
eps = [1e-3, 1e-2, 1e-1, 1.0, 10.0]
model = build_model({"implementation": "vit", "model_name": name, "pretrained": False, "in21k": True, "finetuning": True, "n_classes": 10}, device="cuda").eval()
est = PlasticityEstimator(model)
x, n = torch.randn(P, 3, 224, 224, device="cuda"), torch.randn(P, 3, 224, 224, device="cuda")
for _ in range(3):
    est.sweep_squared_distances(x, n, eps)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    est.sweep_squared_distances(x, n, eps)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
print(f"ViT-{name}: {P} images x {len(eps)} eps: {ms:.3f} ms/chunk -> {P * len(eps) / ms * 1e3:.0f} pairs/s")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        est.sweep_squared_distances(x, n, eps)
    torch.cuda.synchronize()
rows = [(ev.key, ev.device_time_total / 2e3, ev.count // 2) for ev in prof.key_averages() if ev.device_time_total > 0 and ev.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device kernel time per chunk: {tot:.3f} ms (wall {ms:.3f} ms)")
for k, t, c in rows[:16]:
    print(f"{t:9.3f} ms {100*t/tot:5.1f}%  x{c:<5d} {k[:120]}")
