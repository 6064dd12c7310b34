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


class FusedSGD(torch.optim.Optimizer):
    """torch.optim.SGD(momentum, dampening 0, no nesterov) + clip_grad_norm_ as two kernel launches over a flat arena.

    Every trainable parameter's ``.grad`` is a persistent view into one fp32 arena (this optimizer's own, or the
    ``DataParallel`` wrapper's all-reduce buckets when ``grad_arena`` is given), zeroed by a single memset in
    ``zero_grad``; the backward kernels accumulate into it in place (``ops.grad_target``). ``step(max_norm)`` computes
    the global gradient norm, the clip coefficient of ``torch.nn.utils.clip_grad_norm_`` (train.py:277-278), the
    momentum update and the parameter update (optim.py:76-82) in one pass: 20 bytes of HBM traffic per trainable
    element instead of the ~10 foreach / norm passes it replaces. Same arithmetic, element by element, as the
    reference's sequence; the parameters themselves stay separate fp32 tensors (the ``state_dict`` is unchanged).
    Frozen parameters (``requires_grad == False`` at construction, i.e. after ``freeze_model``) are ignored.
    """

    CHUNK = 65536

    def __init__(self, params, lr: float = 1e-3, momentum: float = 0.0, weight_decay: float = 0.0, grad_arena=None):
        # the remaining keys are torch.optim.SGD's own group defaults: a state dict of this optimizer loads into the unfused one
        super().__init__(list(params), dict(lr=lr, momentum=momentum, weight_decay=weight_decay, dampening=0.0, nesterov=False, maximize=False,
                                            foreach=None, differentiable=False, fused=None))
        if len(self.param_groups) != 1:
            raise ValueError("FusedSGD supports a single parameter group")
        self.trainable = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not self.trainable:
            raise ValueError("FusedSGD: no trainable parameter")
        for p in self.trainable:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise TypeError("FusedSGD needs contiguous fp32 CUDA parameters (there is no CPU path)")
            if p.data_ptr() % 16 != 0:
                raise ValueError("FusedSGD: every parameter must be 16-byte aligned (the update kernel uses 128-bit accesses)")
        dev = self.trainable[0].device
        pad4 = lambda n: (n + 3) // 4 * 4
        if grad_arena is not None:  # share DataParallel's flat all-reduce arena
            self.arena, offsets = grad_arena.arena, grad_arena.arena_offset
            self.offset = {p: offsets[p] for p in self.trainable}
        else:
            self.offset, off = {}, 0
            for p in self.trainable:
                self.offset[p] = off
                off += pad4(p.numel())
            self.arena = torch.zeros(off, device=dev, dtype=torch.float32)
        self.momentum_arena = torch.zeros_like(self.arena) if momentum else None
        # bf16 shadows (the GEMM operands, ops.shadow_bf16) of the weight matrices are rewritten by the update kernel itself:
        # no cast launch per matrix and no second read of the parameters after the step
        from . import ops

        self._shadowed = [p for p in self.trainable if p.dim() >= 2 and p.shape[0] > 1]
        shadow_ptr = {p: ops.shadow_bf16(p).data_ptr() for p in self._shadowed}
        rows = []
        for p in self.trainable:
            sh = shadow_ptr.get(p, 0)
            for c0 in range(0, p.numel(), self.CHUNK):
                rows.append((p.data_ptr() + 4 * c0, self.offset[p] + c0, min(self.CHUNK, p.numel() - c0), sh + 2 * c0 if sh else 0))
        self.table = torch.tensor(rows, dtype=torch.int64).to(dev)  # {ptr, arena offset, count | pad, shadow ptr} = 32 bytes per chunk
        self.n_chunks = len(rows)
        self._ptrs = [p.data_ptr() for p in self.trainable]
        self.partials = torch.zeros(1024, device=dev, dtype=torch.float32)  # per-block sums of squares (fixed-order reduction)
        self.grad_norm = torch.zeros(1, device=dev, dtype=torch.float32)
        # hyper-parameters live in device memory (the kernels read them there), so a step captured in a CUDA graph follows
        # an LR schedule / Adam's bias corrections; re-uploaded (from pageable memory: staged at call time) only on change
        self.hyper = torch.zeros(8, device=dev, dtype=torch.float32)
        self._hyper_sent = None
        self._steps = 0
        # torch.optim.AdamW's per-parameter state["step"]: one CPU fp32 scalar PER parameter, as torch keeps them (a single
        # shared tensor would be incremented once per parameter by torch's foreach step after a hand-over)
        self._step_tensors = [torch.tensor(0.0, dtype=torch.float32) for _ in self.trainable]
        self._register_state()
        self.zero_grad()

    def _arena_view(self, arena, p):
        off = self.offset[p]
        return arena[off : off + p.numel()].view_as(p)

    def _state_arenas(self) -> dict:
        """state-dict key -> arena, in torch.optim.SGD's vocabulary, so checkpoints interchange with the unfused optimizer"""
        return {"momentum_buffer": self.momentum_arena} if self.momentum_arena is not None else {}

    def _register_state(self) -> None:
        """Expose the arenas through ``self.state`` as per-parameter views (what ``state_dict()`` / the reference's
        ``Checkpointer`` — ``torch.distributed.checkpoint.state_dict.get_state_dict`` — serialise)."""
        for p, st in zip(self.trainable, self._step_tensors):
            for key, arena in self._state_arenas().items():
                self.state[p][key] = self._arena_view(arena, p)
            if self._has_step_state():
                self.state[p]["step"] = st
        self.param_groups[0]["fused_steps"] = self._steps

    def _has_step_state(self) -> bool:
        """torch.optim.SGD keeps no step count in its state; AdamW does (bias correction)."""
        return False

    def load_state_dict(self, state_dict) -> None:
        """Accepts state dicts of this class AND of the unfused torch optimizer it replaces (torch.optim.SGD /
        torch.optim.AdamW: same per-parameter keys). Loaded tensors are copied INTO the arenas (torch would otherwise
        re-point ``self.state`` at fresh tensors the kernels never read); a parameter without loaded state (torch.optim.SGD
        before its first step) gets zeros. The step count (Adam bias correction) comes from ``fused_steps`` or, for a torch
        AdamW checkpoint, from the per-parameter ``state["step"]``."""
        super().load_state_dict(state_dict)
        steps = self.param_groups[0].get("fused_steps")
        for p in self.trainable:
            st = self.state.get(p, {})
            for key, arena in self._state_arenas().items():
                view = self._arena_view(arena, p)
                loaded = st.get(key)
                if loaded is None:
                    view.zero_()
                elif loaded.data_ptr() != view.data_ptr():
                    view.copy_(loaded)
                self.state[p][key] = view
            if steps is None and "step" in st:
                steps = int(float(st["step"]))
        self._steps = int(steps or 0)
        self._step_tensors = [torch.tensor(float(self._steps), dtype=torch.float32) for _ in self.trainable]
        self._register_state()

    def _slot(self, p):
        off = self.offset[p]
        return self.arena[off : off + p.numel()].view_as(p)

    def zero_grad(self, set_to_none: bool = True) -> None:  # noqa: ARG002 - gradients stay views of the arena
        self.arena.zero_()
        for p in self.trainable:
            g = p.grad
            if g is None or g.data_ptr() != self.arena.data_ptr() + 4 * self.offset[p]:
                p.grad = self._slot(p)

    def _hyper_values(self, max_norm: float) -> list[float]:
        group = self.param_groups[0]
        return [group["lr"], group["momentum"], group["weight_decay"], max_norm]

    def _sync_hyper(self, max_norm: float | None) -> None:
        """Upload the hyper-parameters of the NEXT step if they changed (never inside a graph capture)."""
        vals = self._hyper_values(float("inf") if max_norm is None else float(max_norm))
        if vals != self._hyper_sent:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedSGD: hyper-parameters changed inside a CUDA-graph capture; call _sync_hyper before capturing")
            self.hyper[: len(vals)].copy_(torch.tensor(vals, dtype=torch.float32))
            self._hyper_sent = vals

    def _finish_step(self) -> None:
        """Host-side bookkeeping of one executed step (also called after every replay of a captured step)."""
        from . import ops

        self._steps += 1
        if self._has_step_state():
            torch._foreach_add_(self._step_tensors, 1.0)
        self.param_groups[0]["fused_steps"] = self._steps
        for p in self.trainable:  # the kernel wrote through raw pointers: let version-keyed caches see it
            torch.autograd.graph.increment_version(p)
        for p in self._shadowed:  # ... except the bf16 shadows, which the same kernel rewrote
            ops.mark_shadow_fresh(p)

    @torch.no_grad()
    def step(self, closure=None, max_norm: float | None = None):
        """One update; ``max_norm`` = gradient-clipping threshold (None = no clipping). Returns the pre-clip gradient norm
        as a 0-dim device tensor (what train.py logs as grad_norm)."""
        from . import _lib as L

        assert closure is None
        capturing = torch.cuda.is_current_stream_capturing()
        for p, ptr in zip(self.trainable, self._ptrs):
            if p.data_ptr() != ptr:
                raise RuntimeError("FusedSGD: a parameter was re-allocated after the optimizer was built")
            g = p.grad
            if g is None:
                continue  # no gradient this step: its arena slot is zero
            if g.data_ptr() != self.arena.data_ptr() + 4 * self.offset[p]:  # someone replaced .grad: fold it in
                self._slot(p).add_(g)
                p.grad = self._slot(p)
        if not capturing:
            self._sync_hyper(max_norm)
        L.sumsq_partials_f32(self.arena, self.partials)
        self._update(float("inf") if max_norm is None else max_norm)
        if not capturing:  # a capture executes nothing: GraphedTrainStep books every replay
            self._finish_step()
        return self.grad_norm[0]

    def _update(self, max_norm: float) -> None:
        from . import _lib as L

        group = self.param_groups[0]
        # first_step = False always: the momentum arena starts at zero, and momentum * 0 + g == g is exactly torch's
        # first-step "buf = grad"; a special case here would overwrite momentum buffers loaded from a checkpoint
        L.sgd_momentum_clip_step(self.table, self.n_chunks, self.arena, self.momentum_arena, self.partials, self.grad_norm, max_norm,
                                 group["lr"], group["momentum"], group["weight_decay"], False, hyper=self.hyper)


class FusedAdamW(FusedSGD):
    """torch.optim.AdamW (betas (0.9, 0.999), eps 1e-8, amsgrad off — what optim.py:83-88 builds) + clip_grad_norm_ over
    the same flat gradient arena as :class:`FusedSGD`: norm, clip coefficient, decoupled weight decay, both moment updates
    and the parameter update in one pass (28 bytes of HBM traffic per trainable element)."""

    def __init__(self, params, lr: float = 1e-3, weight_decay: float = 1e-2, betas=(0.9, 0.999), eps: float = 1e-8, grad_arena=None):
        super().__init__(params, lr=lr, momentum=0.0, weight_decay=weight_decay, grad_arena=grad_arena)
        self.param_groups[0].update(betas=tuple(betas), eps=eps, amsgrad=False, capturable=False, decoupled_weight_decay=True)  # torch.optim.AdamW's keys
        self.exp_avg = torch.zeros_like(self.arena)
        self.exp_avg_sq = torch.zeros_like(self.arena)
        self._register_state()

    def _state_arenas(self) -> dict:
        if not hasattr(self, "exp_avg"):  # (called once from the base constructor, before the moment arenas exist)
            return {}
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}

    def _has_step_state(self) -> bool:
        return hasattr(self, "exp_avg")

    def _hyper_values(self, max_norm: float) -> list[float]:
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        t = self._steps + 1
        return [group["lr"], group["weight_decay"], max_norm, group["lr"] / (1.0 - b1**t), 1.0 / math.sqrt(1.0 - b2**t)]

    def _update(self, max_norm: float) -> None:
        from . import _lib as L

        group = self.param_groups[0]
        b1, b2 = group["betas"]
        t = self._steps + 1
        L.adamw_clip_step(self.table, self.n_chunks, self.arena, self.exp_avg, self.exp_avg_sq, self.partials, self.grad_norm, max_norm,
                          group["lr"], b1, b2, group["eps"], group["weight_decay"], 1.0 - b1**t, 1.0 - b2**t, hyper=self.hyper)


def build_optimizer(model: nn.Module, optimizer: str = "sgd", lr: float = 1e-3, momentum: float = 0.0, weight_decay: float = 0.0, fused: bool = False):
    """Over ALL model.parameters(), frozen ones included (they simply never receive a gradient), optim.py:76-89.
    ``fused=True`` returns :class:`FusedSGD` / :class:`FusedAdamW`; pass the ``DataParallel`` wrapper as ``model`` to share its arena."""
    match optimizer.lower():
        case "sgd":
            if fused:
                from .distributed import DataParallel

                return FusedSGD(model.parameters(), lr=lr, momentum=momentum, weight_decay=weight_decay, grad_arena=model if isinstance(model, DataParallel) else None)
            return torch.optim.SGD(model.parameters(), lr=lr, weight_decay=weight_decay, momentum=momentum)
        case "adamw":
            if fused:
                from .distributed import DataParallel

                return FusedAdamW(model.parameters(), lr=lr, weight_decay=weight_decay, grad_arena=model if isinstance(model, DataParallel) else None)
            return torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay)
        case _:
            raise ValueError(f"Unknown optimizer '{optimizer}'. Choose between 'adamw' and 'sgd'.")


def lr_factor(step: int, scheduler: str, n_steps: int, warmup: int = 2000, min_factor: float = 0.0, decay_fraction: float = 0.1,
              cycle_length: float = 1.0) -> float:
    """constant / linear / cosine / wsd multipliers with warm-up (optim.py:119-265; defaults of SchedulerConfig, :113-116).
    wsd (warm-up, stable, decay; optim.py:200-265): cycles of ``n_steps * cycle_length`` steps whose last ``decay_fraction``
    decays as 1 / (t / min_factor + 1 - t)."""
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
        case "wsd":
            cycle = int(n_steps * cycle_length)
            end = cycle * (step // cycle + 1)  # last step of the current cycle
            if step == n_steps:  # the final step closes the last cycle instead of opening a new one
                end = n_steps
            decay = int(end * decay_fraction)
            if step < warmup:
                return float(step) / warmup
            if step <= end - decay:
                return 1.0
            t = (step - (end - decay)) / decay
            return 1.0 / (t / min_factor + (1.0 - t))
        case _:
            raise ValueError(f"Unknown scheduler '{scheduler}'. Choose between 'constant', 'linear', 'cosine' and 'wsd'.")


def build_scheduler(optimizer, scheduler: str, n_steps: int, min_factor: float = 0.0, warmup: int = 2000, decay_fraction: float = 0.1,
                    cycle_length: float = 1.0):
    return torch.optim.lr_scheduler.LambdaLR(optimizer, lambda s: lr_factor(s, scheduler, n_steps, warmup, min_factor, decay_fraction, cycle_length))


def train_step(model, optimizer, batches, grad_clip: float | None, scheduler=None, after_backward=None):
    """One optimisation step over ``batches`` = list of (x, y) micro-batches (len = grad_acc_steps).

    ``after_backward`` (optional callable) runs after the last backward and before clipping — the data-parallel
    wrapper uses it to finish the gradient all-reduce. Returns (last micro-batch loss, pre-clip grad norm), both
    0-dim device tensors (no host sync here; the reference syncs only every logging period, train.py:297-321)."""
    acc = len(batches)
    loss = None
    # gradient accumulation under data parallelism: only the LAST micro-batch's backward may launch the bucket
    # all-reduces (DDP.no_sync semantics); earlier micro-batches accumulate locally into the arena. Without this the
    # overlapped buckets would be reduced during the first backward and later micro-batches would add un-reduced
    # gradients on top of (and racing with) the collective.
    syncer = model if hasattr(model, "require_grad_sync") else None
    # weight / bias gradients that accumulate in place into the arena run on a side stream during backward (their tail
    # waves and the activation-gradient chain's fill each other's idle SMs); joined when the block is left
    from . import ops

    with ops.wgrad_overlap(isinstance(optimizer, FusedSGD)):
        for i, (x, y) in enumerate(batches):
            if syncer is not None:
                syncer.require_grad_sync = i == acc - 1
            preds = model(x)
            loss = F.cross_entropy(preds, y) / acc
            loss.backward()
    if after_backward is not None:
        after_backward()
    if isinstance(optimizer, FusedSGD):
        grad_norm = optimizer.step(max_norm=grad_clip)  # norm + clip + momentum + update in one pass over the arena
    else:
        max_norm = grad_clip if grad_clip is not None else float("inf")
        grad_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
        optimizer.step()
    if scheduler is not None:
        scheduler.step()
    optimizer.zero_grad()
    return loss.detach() * acc, grad_norm


class GraphedTrainStep:
    """``train_step`` for fixed-shape batches, captured ONCE in a CUDA graph and replayed: one graph launch per optimisation
    step instead of ~300 kernel launches issued from Python (forward, loss, backward, the data-parallel bucket all-reduces,
    clip + optimizer step and the gradient-arena reset are all inside the graph).

    The first call runs the step eagerly (it also performs every one-time initialisation a capture must not contain), the
    second call captures and replays, later calls copy the batch into the static input buffers and replay. Requires the
    fused arena optimizers (:class:`FusedSGD` / :class:`FusedAdamW`): their hyper-parameters are read from device memory, so
    an LR schedule (``scheduler.step()`` runs on the host after every replay) or Adam's bias corrections need no re-capture.
    Returns (loss, pre-clip grad norm) like ``train_step``: 0-dim device tensors that the NEXT call overwrites.
    """

    def __init__(self, model, optimizer, grad_clip: float | None, scheduler=None, after_backward=None, grad_acc_steps: int = 1):
        if not isinstance(optimizer, FusedSGD):
            raise TypeError("GraphedTrainStep needs a fused arena optimizer (build_optimizer(..., fused=True))")
        self.model, self.optimizer, self.grad_clip, self.scheduler, self.after_backward = model, optimizer, grad_clip, scheduler, after_backward
        self.acc = grad_acc_steps
        self.graph = None
        self.static = None  # [(x, y)] static input buffers
        self.out = None
        self._calls = 0
        self._sig = None
        self.launches_per_step = None

    def _signature(self, batches):
        return tuple((tuple(x.shape), x.dtype, tuple(y.shape), y.dtype) for x, y in batches) + (self.grad_clip,)

    def _capture(self, batches):
        from . import _lib as L

        dev = batches[0][0].device
        self.static = [(torch.empty_like(x), torch.empty_like(y)) for x, y in batches]
        for (sx, sy), (x, y) in zip(self.static, batches):
            sx.copy_(x)
            sy.copy_(y)
        # every weight's bf16 shadow is rewritten by the optimizer kernel inside the graph; anything still stale now (e.g. a
        # frozen matrix touched by load_state_dict) is cast here, outside, once
        self.optimizer._sync_hyper(self.grad_clip)
        torch.cuda.synchronize(dev)
        before = L.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        # captured on a high-priority stream: the kernel nodes of the step's main chain keep that priority, so they get the
        # SMs before the weight-gradient kernels of the (default-priority) side stream whenever both are runnable
        import os

        hp = os.environ.get("VB_GRAPH_PRIORITY", "1") != "0"
        with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(device=dev, priority=-1) if hp else None):
            self.out = train_step(self.model, self.optimizer, self.static, self.grad_clip, scheduler=None, after_backward=self.after_backward)
        self.launches_per_step = L.launch_count() - before

    def __call__(self, batches):
        if len(batches) != self.acc:
            raise ValueError(f"GraphedTrainStep was built for {self.acc} micro-batch(es), got {len(batches)}")
        self._calls += 1
        sig = self._signature(batches)
        if self._calls == 1 or (self.graph is None and self._sig != sig):
            self._sig = sig
            return train_step(self.model, self.optimizer, batches, self.grad_clip, self.scheduler, self.after_backward)
        if self._sig != sig:
            raise ValueError("GraphedTrainStep: batch shapes / dtypes changed after the capture")
        if self.graph is None:
            self._capture(batches)
        else:
            for (sx, sy), (x, y) in zip(self.static, batches):
                if x.data_ptr() != sx.data_ptr():
                    sx.copy_(x, non_blocking=True)
                if y.data_ptr() != sy.data_ptr():
                    sy.copy_(y, non_blocking=True)
        self.optimizer._sync_hyper(self.grad_clip)
        self.graph.replay()
        self.optimizer._finish_step()
        if self.scheduler is not None:
            self.scheduler.step()
        return self.out
