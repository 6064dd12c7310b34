"""get_probes (reference contract: 96 full taps to the host) vs get_pooled_probes (pooled + normalised on the device)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import build_model  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = build_model({"implementation": "vit", "model_name": "base", "pretrained": False, "in21k": True, "finetuning": True, "n_classes": 10}, device="cuda").eval()
x = torch.randn(N, 3, 224, 224, device="cuda")
for name, fn in (("get_probes + host pooling", lambda: {k: v[:, 0, :] for k, v in model.get_probes(x).items()}),
                 ("get_pooled_probes", lambda: model.get_pooled_probes(x, cls_pooling=True))):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name:28s} {dt * 1e3:9.1f} ms per batch of {N} -> {N / dt:8.1f} img/s")
