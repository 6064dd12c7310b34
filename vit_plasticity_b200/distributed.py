"""Data-parallel finetuning: torchrun environment contract + bucketed, backward-overlapped gradient all-reduce.

Reference: src/vitef/distributed.py. Its env helpers (48-89), ``ComputingManagerConfig`` (140-159) and the
``ComputingManager`` context manager (162-250: ``init_process_group(backend="cpu:gloo,cuda:nccl")`` at 203, device =
``cuda:{LOCAL_RANK}`` at 205, ``DDP(model)`` at 240) are mirrored for the DP branch. TP (232) and FSDP (237) are
self-described work in progress there, have no caller and no plan, and are out of scope: ``tp > 1`` raises.

``DataParallel`` is the DDP equivalent for this package. Built AFTER ``freeze_model`` so that frozen parameters are
excluded, it packs the trainable parameters' gradients into flat fp32 buckets in reverse
registration order (the order backward produces them), and, from post-accumulate-grad hooks, launches one asynchronous
all-reduce per bucket (NCCL over NVLink 5 / NVSwitch on GPUs) as soon as the bucket is complete, so communication
overlaps the rest of backward. ``p.grad`` is re-pointed at the bucket slot, so there is no copy-out, and
``clip_grad_norm_`` / the optimizer see the averaged gradient (identical on every rank).
"""

from __future__ import annotations

import logging
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist
import torch.nn as nn

logger = logging.getLogger("vitef")


def is_torchrun_job() -> bool:
    return os.environ.get("LOCAL_RANK") is not None


def is_distributed_job() -> bool:
    return is_torchrun_job()


def get_rank() -> int:
    return int(os.environ["RANK"]) if is_torchrun_job() else 0


def get_local_rank() -> int:
    return int(os.environ["LOCAL_RANK"]) if is_torchrun_job() else 0


def get_world_size() -> int:
    return int(os.environ["WORLD_SIZE"]) if is_torchrun_job() else 1


def is_master_process() -> bool:
    return get_rank() == 0


@dataclass
class ComputingManagerConfig:
    device: str = "cuda" if torch.cuda.is_available() else "cpu"
    backend: str = "cpu:gloo,cuda:nccl"
    dp: int = 0
    tp: int = 1
    bucket_mb: int = 256

    def __post_init__(self) -> None:
        if not self.dp:
            self.dp = get_world_size() // self.tp


class DataParallel(nn.Module):
    """Gradient-averaging wrapper; ``forward`` delegates to the wrapped module."""

    def __init__(self, module: nn.Module, process_group=None, bucket_mb: int = 256, overlap: bool | None = None):
        super().__init__()
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("DataParallel: the module has no trainable parameter")
        # overlap=True: one all-reduce per bucket, launched from the gradient hooks while backward still runs;
        # overlap=False: ONE all-reduce of the whole arena after backward (NCCL's CTAs then never take SMs away from the
        # persistent GEMM / attention kernels, which are sized to all 148 SMs). Environment overrides for measurement:
        # VB_DP_OVERLAP=0/1, VB_DP_BUCKET_MB=<int>.
        if overlap is None:
            overlap = os.environ.get("VB_DP_OVERLAP", "1") != "0"
        self.overlap = bool(overlap)
        bucket_mb = int(os.environ.get("VB_DP_BUCKET_MB", bucket_mb))
        cap = max(1, int(bucket_mb * 1024 * 1024 // 4))
        # One flat fp32 arena holds every trainable gradient (slots 16-byte aligned, in reverse registration order: the
        # order backward produces them); the buckets are consecutive slices of it. finetune.FusedSGD steps straight from
        # this arena, so the averaged gradients are never copied out.
        pad4 = lambda n: (n + 3) // 4 * 4
        order = list(reversed(params))
        for p in order:
            if p.dtype != torch.float32:
                raise TypeError("DataParallel expects fp32 parameters")
        self.arena = torch.zeros(sum(pad4(p.numel()) for p in order), device=order[0].device, dtype=torch.float32)
        self.arena_offset: dict[torch.Tensor, int] = {}
        self.buckets: list[torch.Tensor] = []
        self._slot: dict[torch.Tensor, tuple[int, torch.Tensor]] = {}
        self._expected: list[int] = []
        off = b_start = 0
        n_in_bucket = 0
        for p in order:
            if n_in_bucket and off - b_start + p.numel() > cap:
                self.buckets.append(self.arena[b_start:off])
                self._expected.append(n_in_bucket)
                b_start, n_in_bucket = off, 0
            self.arena_offset[p] = off
            self._slot[p] = (len(self.buckets), self.arena[off : off + p.numel()].view_as(p))
            off += pad4(p.numel())
            n_in_bucket += 1
        self.buckets.append(self.arena[b_start:off])
        self._expected.append(n_in_bucket)
        self._used = off
        # the mean is taken by NCCL (ncclAvg) where the backend has it; gloo (the CPU tests) sums and divides afterwards
        backend = dist.get_backend(process_group) if dist.is_initialized() else "none"
        self._op = dist.ReduceOp.AVG if backend == "nccl" else dist.ReduceOp.SUM
        self._ready = [0] * len(self.buckets)
        self._seen: set = set()
        self._handles: list = []
        self._launched = [False] * len(self.buckets)
        self.require_grad_sync = True
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        # kernels that accumulate straight into a persistent p.grad (ops.grad_target) bypass autograd's AccumulateGrad
        # and therefore its hooks: they report completed gradients through this registry instead
        from . import ops

        ops.register_grad_ready_hook(self)
        # replicas must start identical: broadcast rank 0's parameters and buffers
        if self.world > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
                # `.data` has its own version counter: bump the tensor's, or version-keyed caches (the bf16 shadow weights
                # of ops.shadow_bf16, PlasticityEstimator) built by an earlier forward would keep the pre-broadcast values
                torch.autograd.graph.increment_version(t)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.module, name)

    # ------------------------------------------------------------------------------------------
    def _on_grad(self, p: torch.Tensor) -> None:
        if p not in self._slot:
            return
        b, slot = self._slot[p]
        if p.grad.data_ptr() != slot.data_ptr():
            slot.copy_(p.grad)
            p.grad = slot  # the optimizer / clip read (and zero_grad drops) the bucket view; no copy-out later
        if not self.require_grad_sync:
            return
        # Idempotent per step: a gradient accumulated in place is reported by ops.grad_done while its kernel is being
        # enqueued, and autograd still runs the post-accumulate hook afterwards (with an undefined incoming gradient).
        if p in self._seen:
            return
        self._seen.add(p)
        self._ready[b] += 1
        if self.overlap and self._ready[b] == self._expected[b] and not self._launched[b]:
            self._launch(b)

    def _launch(self, b: int) -> None:
        self._launched[b] = True
        from . import ops

        ops.side_join()  # weight gradients enqueued on the side stream land in this bucket too
        if self.world > 1:
            self._handles.append(dist.all_reduce(self.buckets[b], op=self._op, group=self.group, async_op=True))

    def finish_grad_sync(self) -> None:
        """Call after the last backward of a step: flushes incomplete buckets (parameters that received no gradient
        contribute zeros), waits for the collectives and turns sums into means."""
        if not self.require_grad_sync:
            return
        for p, (b, slot) in self._slot.items():
            if p.grad is None:
                slot.zero_()
                p.grad = slot
        if not self.overlap and self.world > 1 and not any(self._launched):
            self._launched = [True] * len(self.buckets)
            self._handles.append(dist.all_reduce(self.arena[: self._used], op=self._op, group=self.group, async_op=True))
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._launch(b)
        for h in self._handles:
            h.wait()
        if self.world > 1 and self._op == dist.ReduceOp.SUM:
            for flat in self.buckets:
                flat.div_(self.world)
        self._handles.clear()
        self._seen.clear()
        self._ready = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)

    def grad_bytes(self) -> int:
        return sum(b.numel() * 4 for b in self.buckets)


class ComputingManager:
    """Context manager with the reference's contract (src/vitef/distributed.py:162-250), DP branch only."""

    def __init__(self, config: ComputingManagerConfig):
        self.config = config
        self.backend = config.backend
        self.device = torch.device(config.device)
        self.tp, self.dp = config.tp, config.dp
        if self.tp != 1:
            raise NotImplementedError("tensor parallelism is a work-in-progress stub in the reference (distributed.py:224-233) and is out of scope")
        nb = get_world_size()
        assert self.device.type == "cpu" or self.dp * self.tp == nb, f"DP * TP must equal the number of GPUs {self.tp} * {self.dp} != {nb}"
        os.environ.setdefault("OMP_NUM_THREADS", "1")  # OsEnvironment default, distributed.py:101

    def __enter__(self):
        if not is_distributed_job():
            self.dp = self.tp = 1
            return self
        if self.device.type == "cuda":
            self.device = torch.device(f"cuda:{get_local_rank()}")
            torch.cuda.set_device(self.device)
        dist.init_process_group(backend=self.backend, rank=get_rank(), world_size=get_world_size())
        return self

    def build_model(self, model: nn.Module) -> nn.Module:
        model = model.to(device=self.device)
        if self.dp > 1:
            model = DataParallel(model, bucket_mb=self.config.bucket_mb)
        return model

    def __exit__(self, exc, value, tb):
        if is_distributed_job() and dist.is_initialized():
            dist.destroy_process_group()


def build_manager(config: dict) -> ComputingManager:
    known = {k: v for k, v in config.items() if k in ComputingManagerConfig.__dataclass_fields__}
    return ComputingManager(ComputingManagerConfig(**known))


def get_raw_model(model: nn.Module) -> nn.Module:
    return get_raw_model(model.module) if isinstance(model, DataParallel) else model


def shard_range(n_items: int, rank: int | None = None, world: int | None = None) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of ``n_items`` independent units (input pairs of the plasticity sweep) for this
    rank; no collective is needed on the data path."""
    rank = get_rank() if rank is None else rank
    world = get_world_size() if world is None else world
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)
