"""Kernel-time breakdown of one plasticity-estimator call (CUPTI via torch.profiler)."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import build_model  # noqa: E402
from vit_plasticity_b200.plasticity import PlasticityEstimator  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
name = sys.argv[2] if len(sys.argv) > 2 else "base"
model = build_model({"implementation": "vit", "model_name": name, "pretrained": False, "in21k": True, "finetuning": True, "n_classes": 10}, device="cuda").eval()
est = PlasticityEstimator(model)
x1, x2 = torch.randn(P, 3, 224, 224, device="cuda"), torch.randn(P, 3, 224, 224, device="cuda")
for _ in range(3):
    est.squared_distances(x1, x2)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    est.squared_distances(x1, x2)
e.record()
torch.cuda.synchronize()
print(f"{P} pairs: {s.elapsed_time(e) / 10:.3f} ms/call -> {P / (s.elapsed_time(e) / 10) * 1e3:.0f} pairs/s")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        est.squared_distances(x1, x2)
    torch.cuda.synchronize()
rows = [(ev.key, ev.device_time_total / 2e3, ev.count // 2) for ev in prof.key_averages() if ev.device_time_total > 0 and ev.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device kernel time per call: {tot:.3f} ms")
for k, ms, n in rows[:14]:
    print(f"{ms:9.3f} ms {100*ms/tot:5.1f}%  x{n:<5d} {k[:110]}")
