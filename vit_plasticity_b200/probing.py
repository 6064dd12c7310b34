"""Feature extraction for linear probing, sharded over the ranks (third row).

Drop-in for ``get_embeddings`` of apps/vit/linear_probing.py:58-116: for every batch of the loader the 8 taps per block
of ``get_probes`` are pooled (CLS row or token mean) and, after concatenation, L2-normalised per row. Here the pooling and
the normalisation run on the device inside ``get_pooled_probes`` (``vb_pool_tokens``), so (N, D) rows instead of
(N, 197, D) tensors leave the GPU, and under ``torch.distributed`` every rank takes a contiguous shard of the batches'
samples — each sample is an independent unit — with ONE gather of the pooled rows (and labels) to rank 0 at the end.
"""

from __future__ import annotations

import numpy as np
import torch

from .distributed import shard_range


def _inner(model):
    return model.model if hasattr(model, "model") and hasattr(model.model, "blocks") else model


@torch.inference_mode()
def get_embeddings(model, loader, cls_pooling: bool, device=None, rank: int | None = None, world: int | None = None, group=None):
    """Returns ({tap: (N, D) float32 ndarray, rows L2-normalised}, (N,) labels) on rank 0 — the reference's return value —
    and (None, None) on the other ranks. ``loader`` yields (images, labels); every rank iterates the same loader and keeps
    the samples [lo, hi) of each batch that fall into its shard (``distributed.shard_range`` of the batch)."""
    import torch.distributed as dist

    net = _inner(model)
    dev = device if device is not None else next(net.parameters()).device
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if world is None:
        world = dist.get_world_size(group) if distributed else 1
    if rank is None:
        rank = dist.get_rank(group) if distributed else 0
    model.eval()
    feats: dict[str, list[torch.Tensor]] = {}
    labels: list[torch.Tensor] = []
    counts: list[int] = []  # samples per batch (the gather re-interleaves the ranks' shards batch by batch)
    for x_batch, y_batch in loader:
        n = x_batch.shape[0]
        lo, hi = shard_range(n, rank, world)
        counts.append(n)
        if hi > lo:
            pooled = net.get_pooled_probes(x_batch[lo:hi].to(dev, non_blocking=True), cls_pooling=cls_pooling, normalize=True)
            for key, val in pooled.items():
                feats.setdefault(key, []).append(val)
            labels.append(torch.as_tensor(y_batch[lo:hi]).cpu())
    local = {k: torch.cat(v) for k, v in feats.items()}
    local_labels = torch.cat(labels) if labels else torch.zeros(0, dtype=torch.long)
    if not distributed:
        return {k: v.numpy() for k, v in local.items()}, local_labels.numpy()
    # ---- ONE gather: every rank's rows of every tap, concatenated along the feature axis (same row count per rank and tap) ----
    keys = sorted(local) if local else None
    all_keys = [None] * world
    dist.all_gather_object(all_keys, keys, group=group)
    keys = next(k for k in all_keys if k is not None)
    dims = [None] * world
    dist.all_gather_object(dims, {k: int(local[k].shape[1]) for k in keys} if local else None, group=group)
    dims = next(d for d in dims if d is not None)
    n_local = int(local_labels.shape[0])
    flat = torch.cat([local[k] if local else torch.zeros(0, dims[k]) for k in keys] + [local_labels.float().unsqueeze(1)], 1) if n_local else torch.zeros(0, sum(dims.values()) + 1)
    sizes = [None] * world
    dist.all_gather_object(sizes, n_local, group=group)
    per = max(sizes)
    padded = torch.zeros(per, flat.shape[1])
    padded[:n_local] = flat
    if dist.get_backend(group) == "nccl":
        padded = padded.to(dev)
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, parts, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None, None
    parts = [p.cpu() for p in parts]
    # rank r holds, for every batch, that batch's shard r: rebuild the loader's sample order
    rows, cursor = [], [0] * world
    for n in counts:
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            rows.append(parts[r][cursor[r] : cursor[r] + hi - lo])
            cursor[r] += hi - lo
    table = torch.cat(rows)
    out, col = {}, 0
    for k in keys:
        out[k] = table[:, col : col + dims[k]].numpy()
        col += dims[k]
    return out, table[:, col].long().numpy()
