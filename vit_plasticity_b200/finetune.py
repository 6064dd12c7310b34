"""Finetuning hot loop of apps/vit/train.py on the B200 kernels.

Keeps the reference semantics of one optimisation step (apps/vit/train.py:243-284): accumulate
``F.cross_entropy(model(x), y) / grad_acc_steps`` over ``grad_acc_steps`` micro-batches, then
``clip_grad_norm_`` (max-norm inf when ``grad_clip`` is None, :277-278), ``optimizer.step()``, ``scheduler.step()``,
``optimizer.zero_grad()``. ``freeze_model`` follows apps/vit/utils.py:54-91 (components list what to FREEZE; matching
is by substring on parameter names; the final norm and the head are never frozen). Logging, evaluation and
checkpoint plumbing of the reference app are out of scope and unchanged by this
package; optimizer and LR schedule are built with the reference's recipes (src/vitef/optim.py:76-89, 160-196).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, fields

import torch
import torch.nn as nn
import torch.nn.functional as F

from .models.factory import DEVICE

COMPONENT_WEIGHTS = {
    "emb": ["embedding"],
    "attn_norm": ["attn_norm"],
    "mha": ["attn.qkv_mat", "attn.output"],
    "ffn_norm": ["ffn_norm"],
    "ffn_fc1": ["ffn.fc1"],
    "ffn_fc2": ["ffn.fc2"],
}


def freeze_model(model: nn.Module, components: list[str] | None) -> None:
    """Set ``requires_grad = False`` on the listed components across all blocks."""
    patterns: list[str] = []
    for comp in components or []:  # the reference iterates None unconditionally; every YAML sets []
        patterns.extend(COMPONENT_WEIGHTS[comp])
    inner = model.model if hasattr(model, "model") else model
    if "embedding" in patterns:
        for p in inner.embedding.parameters():
            p.requires_grad = False
    for block in inner.blocks:
        for name, p in block.named_parameters():
            if any(pat in name for pat in patterns):
                p.requires_grad = False


@dataclass
class TrainingConfig:
    """Same fields and defaults as apps/vit/train.py:43-101."""

    model_name: str = "base"
    patch_size: int = 16
    image_dim: tuple = (3, 224, 224)
    components: list | None = None
    dataset_name: str = "cifar10"
    train_size: float = 0.8
    batch_size: int = 512
    val_batch_size: int = 512
    n_steps: int = 10_000
    grad_acc_steps: int = 1
    grad_clip: float | None = None
    eval_period: int = 1000
    optimizer: str = "sgd"
    lr: float = 1e-3
    momentum: float = 0.9
    scheduler: str = "constant"
    min_factor: float = 0
    device: str = DEVICE
    log_dir: str = ""
    overwrite: bool = False
    logging_period: int = 10
    logging_level: str = "INFO"
    seed: int = 42
    utility_period: int = 1000

    def __init__(self, **kwargs):
        for f in fields(self):
            setattr(self, f.name, kwargs.get(f.name, f.default))
        self.__post_init__()

    def __post_init__(self):
        if (self.eval_period <= 0) or (self.eval_period > self.n_steps):
            self.eval_period = self.n_steps
        if self.seed is None:
            self.seed = 42
        if isinstance(self.image_dim, list):
            self.image_dim = tuple(self.image_dim)


def build_optimizer(model: nn.Module, optimizer: str = "sgd", lr: float = 1e-3, momentum: float = 0.0, weight_decay: float = 0.0):
    """Over ALL model.parameters(), frozen ones included (they simply never receive a gradient), optim.py:76-89."""
    match optimizer.lower():
        case "sgd":
            return torch.optim.SGD(model.parameters(), lr=lr, weight_decay=weight_decay, momentum=momentum)
        case "adamw":
            return torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay)
        case _:
            raise ValueError(f"Unknown optimizer '{optimizer}'. Choose between 'adamw' and 'sgd'.")


def lr_factor(step: int, scheduler: str, n_steps: int, warmup: int = 2000, min_factor: float = 0.0) -> float:
    """constant / linear / cosine multipliers with warm-up (optim.py:119-196; default warmup 2000, :113)."""
    match scheduler.lower():
        case "constant":
            return 1.0
        case "cosine" | "linear":
            if step < warmup:
                return float(step) / warmup
            if step <= n_steps:
                s = float(step - warmup) / (n_steps - warmup)
                if scheduler.lower() == "cosine":
                    return min_factor + 0.5 * (1 - min_factor) * (math.cos(math.pi * s) + 1)
                return min_factor + (1 - min_factor) * (1 - s)
            return min_factor
        case _:
            raise ValueError(f"Unknown scheduler '{scheduler}'.")


def build_scheduler(optimizer, scheduler: str, n_steps: int, min_factor: float = 0.0, warmup: int = 2000):
    return torch.optim.lr_scheduler.LambdaLR(optimizer, lambda s: lr_factor(s, scheduler, n_steps, warmup, min_factor))


def train_step(model, optimizer, batches, grad_clip: float | None, scheduler=None, after_backward=None):
    """One optimisation step over ``batches`` = list of (x, y) micro-batches (len = grad_acc_steps).

    ``after_backward`` (optional callable) runs after the last backward and before clipping — the data-parallel
    wrapper uses it to finish the gradient all-reduce. Returns (last micro-batch loss, pre-clip grad norm), both
    0-dim device tensors (no host sync here; the reference syncs only every logging period, train.py:297-321)."""
    acc = len(batches)
    loss = None
    for x, y in batches:
        preds = model(x)
        loss = F.cross_entropy(preds, y) / acc
        loss.backward()
    if after_backward is not None:
        after_backward()
    max_norm = grad_clip if grad_clip is not None else float("inf")
    grad_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    optimizer.zero_grad()
    return loss.detach() * acc, grad_norm
