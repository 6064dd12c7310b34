"""Where the step is NOT running kernels: device-side idle gaps of the finetuning step (torch.profiler / CUPTI timestamps)."""
import sys
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
for _ in range(4):
    train_step(model, opt, [(x, y)], grad_clip=1.0)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    train_step(model, opt, [(x, y)], grad_clip=1.0)
e.record()
torch.cuda.synchronize()
print(f"un-profiled: {s.elapsed_time(e) / 5:.2f} ms / step")
N = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        train_step(model, opt, [(x, y)], grad_clip=1.0)
    torch.cuda.synchronize()
ev = [(k.time_range.start, k.time_range.end, k.name) for k in prof.events() if k.device_type.name == "CUDA" and k.time_range.end > k.time_range.start]
ev.sort()
span = ev[-1][1] - ev[0][0]
busy = 0
gaps = []
cur_end = ev[0][0]
for a, b, name in ev:
    if a > cur_end:
        gaps.append((a - cur_end, prev, name))
    if b > cur_end:
        busy += b - max(a, cur_end)
        cur_end = b
        prev = name
print(f"profiled: span {span / 1e3 / N:.2f} ms / step, busy {busy / 1e3 / N:.2f} ms / step, idle {(span - busy) / 1e3 / N:.2f} ms / step over {len(ev) // N} device ops / step")
import collections

agg = collections.defaultdict(lambda: [0, 0.0])
for g, before, after in gaps:
    k = (before.split("(")[0][-60:], after.split("(")[0][-60:])
    agg[k][0] += 1
    agg[k][1] += g
print("largest idle-gap classes (us per step, count per step, kernel before -> kernel after):")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t / N:9.1f} us  x{n / N:6.1f}  {k[0]}  ->  {k[1]}")
