"""Drives the UNMODIFIED reference (installed under baseline/_ref by baseline/install_ref.sh) on the host cores.

Nothing of this repository's models, kernels or oracle is on this path: the model comes from the reference's
``vitef.models.build_model``, the optimizer from ``vitef.optim.build_optimizer``, component freezing from
``apps.vit.utils.freeze_model``, the estimator from ``ViT.get_decomposition`` + ``apps.vit.analysis.distance``. The
few lines below are the bodies of the reference's own loops (apps/vit/train.py:263-283, apps/vit/analysis.py:216-233),
which live inline in its ``train()`` / ``analysis()`` entry points and cannot be imported on their own.

``fire`` and ``omegaconf`` are absent from the image and used only inside the apps' ``main()``: empty stubs are
injected before import (the reference files themselves are untouched).
"""

from __future__ import annotations

import os
import statistics
import sys
import time
import types
from pathlib import Path

REF = Path(__file__).resolve().parent / "_ref"


def available() -> str | None:
    """None if the reference is installed, else a one-line reason."""
    if not (REF / "vitef").is_dir():
        return f"{REF}/vitef missing: run baseline/install_ref.sh in the build container"
    if not (REF / "apps" / "vit" / "utils.py").is_file():
        return f"{REF}/apps missing: run baseline/install_ref.sh in the build container"
    return None


def _import_reference():
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    for missing in ("fire", "omegaconf"):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except ImportError:
                stub = types.ModuleType(missing)
                stub.OmegaConf = object
                stub.Fire = lambda *a, **k: None
                sys.modules[missing] = stub
    omp = os.environ.get("OMP_NUM_THREADS")
    from apps.vit.analysis import distance  # sets OMP_NUM_THREADS=1 at import (analysis.py:14): restored below
    from apps.vit.utils import freeze_model
    from vitef.models import build_model
    from vitef.optim import build_optimizer

    if omp is None:
        os.environ.pop("OMP_NUM_THREADS", None)
    else:
        os.environ["OMP_NUM_THREADS"] = omp
    return build_model, build_optimizer, freeze_model, distance


def _vit(build_model, model_name: str, device: str):
    import torch

    torch.manual_seed(42)
    cfg = {"implementation": "vit", "model_name": model_name, "pretrained": False, "in21k": True, "finetuning": True, "n_classes": 10}
    return build_model(cfg, device=device)


def finetune(model_name: str, batch: int, steps: int, warmup: int, components, threads: int, device: str = "cpu", autocast: bool = False):
    """The reference's optimisation step (train.py:263-283: forward, F.cross_entropy, backward, clip_grad_norm_ 1.0,
    SGD lr 1e-2 momentum 0.9 of apps/vit/configs/cifar10.yaml, zero_grad) on one synthetic batch per step.
    Returns (img/s from the median step, median seconds per step, trainable parameters)."""
    import torch
    import torch.nn.functional as F
    from torch.nn.utils import clip_grad_norm_

    build_model, build_optimizer, freeze_model, _ = _import_reference()
    if device == "cpu":
        torch.set_num_threads(threads)
    model = _vit(build_model, model_name, device)
    model.train()
    freeze_model(model, list(components))
    optimizer = build_optimizer({"optimizer": "sgd", "lr": 1e-2, "momentum": 0.9}, model)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, 3, 224, 224, generator=g).to(device)
    y = torch.randint(0, 10, (batch,), generator=g).to(device)
    times = []
    for i in range(warmup + steps):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast and device != "cpu"):
            preds = model(x)
            loss = F.cross_entropy(preds, y)
        loss.backward()
        clip_grad_norm_(model.parameters(), 1.0)
        optimizer.step()
        optimizer.zero_grad()
        if device != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = statistics.median(times)
    n_trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    return batch / dt, dt, n_trainable


def plasticity(model_name: str, pairs: int, batch: int, threads: int, device: str = "cpu"):
    """BASELINE.json configs[0]: the reference estimator as apps/vit/analysis.py:203-233 runs it (pin + copy of both
    batches, get_decomposition on each — every component output comes back on the host, architecture.py:385-418 —, then
    per key the re-upload and distance) on ``pairs`` synthetic image pairs, fp32. Returns (pairs/s, seconds)."""
    import torch

    build_model, _, _, distance = _import_reference()
    torch.set_num_threads(threads)
    model = _vit(build_model, model_name, device)
    model.eval()
    g = torch.Generator().manual_seed(0)
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    done = 0
    while done < pairs:
        n = min(batch, pairs - done)
        x1, x2 = torch.randn(n, 3, 224, 224, generator=g), torch.randn(n, 3, 224, 224, generator=g)
        if device != "cpu":
            x1, x2 = x1.pin_memory(), x2.pin_memory()
        x1, x2 = x1.to(device=device, non_blocking=True), x2.to(device=device, non_blocking=True)
        out1, out2 = model.get_decomposition(x1), model.get_decomposition(x2)
        dist = {}
        for key in list(out1.keys()):
            z1, z2 = out1.pop(key).to(device), out2.pop(key).to(device)
            dist[key] = distance(z1, z2, reduction="none").cpu().numpy()
        done += n
    if device != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return pairs / dt, dt


def loss_and_grads_on(device: str, model_name: str, state_dict: dict, x, y, n_classes: int = 10):
    """The unmodified reference ViT on ``device`` (fp32, TF32 off) with the given weights: cross-entropy loss of one batch
    and the gradient of every parameter, {name: tensor}. Used for the bench-scale parity check (batch 512: M = 100 864
    token rows, split-K weight gradients over 100 864 tokens), which no CPU oracle finishes in seconds."""
    import torch
    import torch.nn.functional as F

    build_model, _, _, _ = _import_reference()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = {"implementation": "vit", "model_name": model_name, "pretrained": False, "in21k": True, "finetuning": True, "n_classes": n_classes}
    model = build_model(cfg, device=device)
    model.load_state_dict(state_dict)
    model.train()
    loss = F.cross_entropy(model(x), y)
    loss.backward()
    return float(loss), {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
