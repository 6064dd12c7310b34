"""Fused, on-device per-component plasticity estimator.

Reference path (what this replaces): ``Transformer.get_decomposition`` twice (architecture.py:856-883, every
component output copied to the host), then per key an H2D re-upload and ``distance`` (apps/vit/analysis.py:216-233,
:68), then offline ``ratio = dist[key] / dist["embedding"]`` (apps/plots/analysis.py:97).

Here nothing but the final ``(1 + 5 * n_layers) x N`` table of distances leaves the GPU and no component output or
difference tensor is ever written to HBM. A pair is carried as (a, d): the base input a and the DIFFERENCE d = b - a,
formed in fp32 from the images before anything is rounded to bf16, so the accuracy of every component is independent of
the size of the perturbation (two independent bf16 evaluations of a and b lose the difference once |d| << |a|: 34-53 %
error on the attention component at a relative perturbation of 1e-2):

* every block sees the SAME embedding output e (architecture.py:877-881), so the per-layer weights are concatenated
  and each component family is ONE GEMM over all layers;
* linear components — patch embedding, ``fc1``, ``fc2`` on ``[e,0,0,0]`` (== ``W2[:, :E]``) and the attention output
  projection — satisfy f(b) - f(a) = W d (the bias cancels), so the GEMM runs on the difference and its epilogue
  reduces sum(acc^2) per sample straight out of TMEM (``VB_EPI_SUMSQ``); for a sweep x + eps * n they scale as eps;
* LayerNorm: LN_i(b) - LN_i(a) = gamma_i * (zhat_b - zhat_a); one fp32 kernel evaluates zhat_b - zhat_a from (a, d)
  without cancellation (``vb_layernorm_delta_sqdiff``) and every norm's squared distance is u @ gamma_i^2;
* attention: q/k/v of a and dq/dk/dv = W d come from one concatenated-weight GEMM each; the perturbation-form kernel
  (``vb_attention_perturb_delta_layers``) forms the score, probability and output differences from those small
  operands directly, and the (linear) output projection runs on its result with the sum-of-squares epilogue.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import ops

COMPONENTS = ("attn_norm", "attn", "ffn_norm", "ffn_fc1", "ffn_fc2")  # order of TransformerBlock._decompose


def _inner(model):
    return model.model if hasattr(model, "model") and hasattr(model.model, "blocks") else model


class PlasticityEstimator:
    """Holds the concatenated bf16 weights of a model (rebuilt when a parameter changes)."""

    def __init__(self, model):
        self.net = _inner(model)
        self._key = None

    # ------------------------------------------------------------------------------------------
    def _refresh(self):
        net = self.net
        params = list(net.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key == self._key:
            return
        blocks = list(net.blocks)
        e = net.embedding.pos_dim
        f = blocks[0].ffn.fc1.weight.shape[0]
        if f != 4 * e:
            raise ValueError("the fc2 plasticity tap feeds [x,0,0,0] to fc2 and needs ffn_dim == 4 * emb_dim (architecture.py:414-416)")
        if e % 128 != 0:
            raise ValueError("emb_dim must be a multiple of 128 for the fused sum-of-squares epilogue")
        cat16 = lambda ts: torch.cat([t.detach().reshape(t.shape[0], -1) for t in ts], 0).to(torch.bfloat16).contiguous()
        conv = net.embedding.patching.patching[0]
        self.w_patch = cat16([conv.weight])
        self.b_patch = conv.bias.detach().float().contiguous()
        self.w_fc1 = cat16([b.ffn.fc1.weight for b in blocks])                    # [n_layers * F, E]
        self.w_fc2 = cat16([b.ffn.fc2.weight[:, :e] for b in blocks])             # [n_layers * E, E]
        self.w_qkv = cat16([b.attn.qkv_mat.weight for b in blocks])               # [n_layers * 3E, E]
        self.b_qkv = torch.cat([b.attn.qkv_mat.bias.detach().float() for b in blocks]).contiguous()
        self.w_out = [ops.shadow_bf16(b.attn.output.weight) for b in blocks]
        gammas = [b.attn_norm.weight for b in blocks] + [b.ffn_norm.weight for b in blocks]
        self.gamma_sq = torch.stack([g.detach().float() ** 2 for g in gammas], 1).contiguous()  # [E, 2 * n_layers]
        self.cls = net.embedding.cls_token.detach().float().reshape(-1).contiguous()
        self.pos = net.embedding.pos_emb.detach().float().reshape(-1, e).contiguous()
        self.zeros_cls = torch.zeros_like(self.cls)
        self.zeros_pos = torch.zeros_like(self.pos)
        self.eps = blocks[0].attn_norm.eps
        self.heads = blocks[0].attn.h
        self.e, self.f, self.n_layers = e, f, len(blocks)
        self.patch = net.embedding.patching.patch_size
        self._key = key

    # ------------------------------------------------------------------------------------------
    # the three stages of a pair (a, d): what depends on a only, on d only (linear: scales with |d|), and on both
    # ------------------------------------------------------------------------------------------
    def _base(self, x: torch.Tensor) -> dict:
        """Embedding of the base input (fp32 tokens out of the accumulator + their bf16 copy) and its q/k/v, all layers."""
        n = x.shape[0]
        e, nl = self.e, self.n_layers
        p = L.im2col_patches(x, self.patch)
        rows_p, kp = p.shape
        np_ = rows_p // n
        seq = np_ + 1
        m = n * seq
        po = torch.empty(rows_p, e, device=x.device, dtype=torch.float32)
        L.gemm(p, self.w_patch, m=rows_p, n=e, k=kp, epilogue=L.EPI_F32, bias=self.b_patch, out=po)
        t16, t32 = L.assemble_tokens(None, po, self.cls, self.pos, n, np_, e, want_bf16=True, want_f32=True)
        qkv = torch.empty(m, nl * 3 * e, device=x.device, dtype=torch.bfloat16)
        L.gemm(t16, self.w_qkv, m=m, n=nl * 3 * e, k=e, epilogue=L.EPI_BF16, bias=self.b_qkv, out=qkv)
        return {"n": n, "np": np_, "seq": seq, "m": m, "t16": t16, "t32": t32, "qkv": qkv}

    def _direction(self, xb: torch.Tensor, xa: torch.Tensor | None) -> dict:
        """Everything that is linear in the difference d = xb - xa (xa None: xb IS the difference, e.g. a noise
        direction): embedding distance, fc1 / fc2 distances, the fp32 / bf16 token difference and its q/k/v projections."""
        n = xb.shape[0]
        e, f, nl = self.e, self.f, self.n_layers
        dev = xb.device
        pd = L.im2col_patches(xb, self.patch, xa)  # (xb - xa) subtracted in fp32, then bf16
        rows_p, kp = pd.shape
        np_ = rows_p // n
        seq = np_ + 1
        m = n * seq
        dpo = torch.empty(rows_p, e, device=dev, dtype=torch.float32)
        L.gemm(pd, self.w_patch, m=rows_p, n=e, k=kp, epilogue=L.EPI_F32, out=dpo)  # the bias cancels in the difference
        ss_emb = torch.zeros(n, device=dev, dtype=torch.float32)
        L.rowsumsq_diff_f32(dpo, None, ss_emb, n, np_, e)  # cls / pos rows of the difference are zero
        d16, d32 = L.assemble_tokens(None, dpo, self.zeros_cls, self.zeros_pos, n, np_, e, want_bf16=True, want_f32=True)
        ss_fc1 = torch.zeros(n, nl, device=dev, dtype=torch.float32)
        L.gemm(d16, self.w_fc1, m=m, n=nl * f, k=e, epilogue=L.EPI_SUMSQ, sumsq=ss_fc1, rows_per_sample=seq, cols_per_group=f, n_groups=nl)
        ss_fc2 = torch.zeros(n, nl, device=dev, dtype=torch.float32)
        L.gemm(d16, self.w_fc2, m=m, n=nl * e, k=e, epilogue=L.EPI_SUMSQ, sumsq=ss_fc2, rows_per_sample=seq, cols_per_group=e, n_groups=nl)
        dqkv = torch.empty(m, nl * 3 * e, device=dev, dtype=torch.bfloat16)
        L.gemm(d16, self.w_qkv, m=m, n=nl * 3 * e, k=e, epilogue=L.EPI_BF16, out=dqkv)
        return {"ss_emb": ss_emb, "ss_fc1": ss_fc1, "ss_fc2": ss_fc2, "d32": d32, "dqkv": dqkv}

    def _table(self, base: dict, dirn: dict, scale: float, scratch: dict | None = None) -> torch.Tensor:
        """[1 + 5 n_layers, N] squared distances of the pair (a, a + scale * d)."""
        n, seq, m = base["n"], base["seq"], base["m"]
        e, nl, heads = self.e, self.n_layers, self.heads
        dev = base["t32"].device
        s2 = float(scale) * float(scale)
        # ---- LayerNorms: all 2 * n_layers at once ----
        u = torch.zeros(n, e, device=dev, dtype=torch.float32)
        L.layernorm_delta_sqdiff(base["t32"], dirn["d32"], scale, u, n, seq, e, self.eps)
        ss_ln = u @ self.gamma_sq  # [N, 2 * n_layers]; 2*N*E*2n_layers FLOPs, negligible
        # ---- attention (applied to the un-normalised embedding, architecture.py:405) ----
        ss_attn = torch.zeros(n, nl, device=dev, dtype=torch.float32)
        scratch = scratch if scratch is not None else {}
        if scale == 1.0:
            dqkv = dirn["dqkv"]
        else:
            buf = scratch.get("dqkv")
            if buf is None or buf.shape != dirn["dqkv"].shape:
                buf = scratch["dqkv"] = torch.empty_like(dirn["dqkv"])
            dqkv = L.scale_bf16(dirn["dqkv"], scale, buf)
        if seq <= 208:
            out = scratch.get("delta")
            if out is None or out.shape != (nl, m, e):
                out = scratch["delta"] = torch.empty(nl, m, e, device=dev, dtype=torch.bfloat16)
            delta_all = L.attention_perturb_delta_layers(base["qkv"], dqkv, nl, n, seq, heads, e // heads, out=out)  # one launch, all layers
        else:
            # longer sequences (ViT-H/14: 257 tokens) have no perturbation-form kernel: two independent evaluations on the
            # mma.sync path, accurate for perturbations of the order of the input only
            qkv_b = L.add_bf16(base["qkv"], dqkv)
            delta_all = torch.empty(nl, m, e, device=dev, dtype=torch.bfloat16)
            for i in range(nl):
                qa, qb = base["qkv"][:, i * 3 * e : (i + 1) * 3 * e], qkv_b[:, i * 3 * e : (i + 1) * 3 * e]
                L.attention_pair_delta(qb, qa, delta_all[i], n, seq, heads, e // heads)
        for i in range(nl):
            # sumsq pointer offset by i with n_groups = n_layers writes column i of ss_attn
            L.gemm(delta_all[i], self.w_out[i], m=m, n=e, k=e, epilogue=L.EPI_SUMSQ, sumsq=ss_attn[:, i:], rows_per_sample=seq, cols_per_group=e, n_groups=nl)
        # ---- assemble in the reference's key order; the linear components scale with the perturbation ----
        rows = [dirn["ss_emb"] * s2]
        for i in range(nl):
            rows += [ss_ln[:, i], ss_attn[:, i], ss_ln[:, nl + i], dirn["ss_fc1"][:, i] * s2, dirn["ss_fc2"][:, i] * s2]
        return torch.stack(rows, 0)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def squared_distances(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        """f32 device tensor [1 + 5 * n_layers, N] of SQUARED distances, rows ordered like the reference's keys."""
        if not (x1.is_cuda and x2.is_cuda):
            raise RuntimeError("PlasticityEstimator runs on CUDA only — there is no CPU fallback")
        self._refresh()
        x1, x2 = x1.float().contiguous(), x2.float().contiguous()
        return self._table(self._base(x1), self._direction(x2, x1), 1.0)

    @torch.no_grad()
    def sweep_squared_distances(self, x: torch.Tensor, direction: torch.Tensor, eps_list) -> list[torch.Tensor]:
        """Tables of the pairs (x, x + eps * direction) for every eps: f(x) and everything linear in the direction are
        computed once and shared by the whole grid (apps/plots/loss_landscape.py:180-191 style magnitude sweep)."""
        if not (x.is_cuda and direction.is_cuda):
            raise RuntimeError("PlasticityEstimator runs on CUDA only — there is no CPU fallback")
        self._refresh()
        base = self._base(x.float().contiguous())
        dirn = self._direction(direction.float().contiguous(), None)
        scratch: dict = {}
        return [self._table(base, dirn, float(eps), scratch) for eps in eps_list]

    def keys(self) -> list[str]:
        if self._key is None:
            self._refresh()
        out = ["embedding"]
        for i in range(self.n_layers):
            out += [f"block{i}_{c}" for c in COMPONENTS]
        return out

    @torch.no_grad()
    def pair_distances(self, x1: torch.Tensor, x2: torch.Tensor, max_pairs_per_call: int = 64) -> dict[str, np.ndarray]:
        """Same result as one iteration of the reference analysis loop: {key: (N,) float32 distances}."""
        chunks = []
        for s in range(0, x1.shape[0], max_pairs_per_call):
            chunks.append(self.squared_distances(x1[s : s + max_pairs_per_call], x2[s : s + max_pairs_per_call]))
        table = torch.cat(chunks, 1).sqrt().cpu().numpy()  # the only device -> host copy
        return {k: table[j] for j, k in enumerate(self.keys())}


def gather_tables(local: torch.Tensor, n_total: int, per_rank: int, group=None) -> torch.Tensor | None:
    """Concatenate the ranks' [..., n_local] distance tables along the pair axis on rank 0 (contiguous shards in rank
    order, see ``distributed.shard_range``; every rank pads to ``per_rank`` columns). The only collective of the sweep:
    ONE gather at the end (NCCL on GPUs; the gloo CPU tests gather host tensors)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if dist.get_backend(group) != "nccl":
        local = local.cpu()
    padded = local.new_zeros(*local.shape[:-1], per_rank)
    padded[..., : local.shape[-1]] = local
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded.contiguous(), parts, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None
    # shards are contiguous; only the tail ranks are short (or empty): drop every rank's padding
    table = torch.cat([parts[r][..., : max(0, min(per_rank, n_total - r * per_rank))] for r in range(world)], -1)
    assert table.shape[-1] == n_total, (table.shape, n_total)
    return table


def sweep_noise(n_images: int, image_shape, noise_seed: int, first_image: int, device) -> torch.Tensor:
    """Perturbation directions of images [first_image, first_image + n_images): standard normals drawn ON THE DEVICE from
    Philox4x32-10 keyed by ``noise_seed`` with the image index in the counter (``vb_philox_normal_f32``), so the noise of
    an image depends neither on the batching nor on the sharding over ranks."""
    return L.philox_normal(n_images, tuple(image_shape), noise_seed, first_image, device)


def sweep_local(est: PlasticityEstimator, images, eps_list, noise_seed: int = 0, first_image: int = 0, pairs_per_call: int = 64,
                transform=None, device=None) -> torch.Tensor:
    """Distances of the pairs (x_i, x_i + eps n_i) for the images of ONE shard: f32 device tensor
    [n_eps, 1 + 5 n_layers, n_images]. ``first_image`` is the global index of ``images[0]`` (the noise n_i is keyed by the
    global image index). x + eps n is never formed: the pair is carried as (x, eps n), f(x) and everything linear in n are
    shared by the whole eps grid."""
    dev = device if device is not None else next(est.net.parameters()).device
    eps_list = [float(e) for e in eps_list]
    chunks: list[torch.Tensor] = []
    n = images.shape[0]
    for s0 in range(0, n, pairs_per_call):
        s1 = min(s0 + pairs_per_call, n)
        x = images[s0:s1].to(dev, non_blocking=True)
        x = transform(x) if transform is not None else x.float()
        noise = sweep_noise(s1 - s0, x.shape[1:], noise_seed, first_image + s0, dev)
        chunks.append(torch.stack(est.sweep_squared_distances(x, noise, eps_list), 0).sqrt())  # [n_eps, rows, n_chunk]
    return torch.cat(chunks, 2) if chunks else torch.zeros(len(eps_list), len(est.keys()), 0, device=dev)


def perturbation_sweep(model, images, eps_list, noise_seed: int = 0, pairs_per_call: int = 64, rank: int | None = None,
                       world: int | None = None, group=None, estimator: "PlasticityEstimator | None" = None, stats: dict | None = None,
                       transform=None):
    """Plasticity under input perturbations of growing magnitude (BASELINE.json configs[4]): pairs (x, x + eps * n),
    n ~ N(0, 1) per image from ``noise_seed`` (:func:`sweep_noise`), for every eps in ``eps_list``.

    The N images are sharded contiguously over the ranks (each pair is independent: apps/vit/analysis.py:68 reduces per
    sample); every rank streams its shard from ``images`` (host or device, fp32 NCHW) through :func:`sweep_local`, and only
    the (1 + 5 n_layers) x N_local distance table per eps leaves the GPU. No collective on the data path; rank 0 receives
    {eps: {key: (N,) float32}} from ONE gather at the end, the other ranks get None.
    ``stats`` (optional dict) receives ``gather_s``, the wall time of that gather. ``transform`` (optional callable) maps a
    chunk of ``images`` already on the device to fp32 NCHW — e.g. ``preprocess.DevicePreprocessor(224, "test")`` when
    ``images`` holds the dataset's raw uint8 HWC samples, so only those cross PCIe.
    """
    import time

    import torch.distributed as dist

    from .distributed import shard_range

    est = estimator if estimator is not None else PlasticityEstimator(model)
    n_total = images.shape[0]
    lo, hi = shard_range(n_total, rank, world)
    eps_list = [float(e) for e in eps_list]
    keys = est.keys()
    local = sweep_local(est, images[lo:hi], eps_list, noise_seed, lo, pairs_per_call, transform, next(_inner(model).parameters()).device)
    distributed = dist.is_available() and dist.is_initialized() and (world is None or world == dist.get_world_size(group))
    t0 = time.perf_counter()
    if distributed and dist.get_world_size(group) > 1:
        w = dist.get_world_size(group)
        table = gather_tables(local, n_total, (n_total + w - 1) // w, group)
        if local.is_cuda:
            torch.cuda.synchronize()
    else:
        table = local
    if stats is not None:
        stats["gather_s"] = time.perf_counter() - t0
    if table is None:
        return None
    table = table.cpu().numpy()
    return {e: {k: table[i, j] for j, k in enumerate(keys)} for i, e in enumerate(eps_list)}


def analysis(model, loader1, loader2, n_steps: int, save_dir=None, device=None) -> dict[str, np.ndarray]:
    """Drop-in for the loop of apps/vit/analysis.py:188-248: batch k of ``loader1`` is paired with batch k of
    ``loader2`` (iterables of (images, labels)), the per-sample distances of every component are accumulated with the
    reference's ``update_dict`` semantics and, if ``save_dir`` is given, written to ``save_dir/distances.pkl`` in the
    reference's format (a pickled {key: float32 ndarray of length n_steps * batch}), which apps/plots/analysis.py reads.
    Loaders are re-iterated when exhausted, like ``make_iterable`` (src/vitef/utils.py)."""
    import pickle
    from pathlib import Path

    est = PlasticityEstimator(model)
    dev = device if device is not None else next(_inner(model).parameters()).device

    def forever(loader):
        while True:
            yield from loader

    it1, it2 = forever(loader1), forever(loader2)
    distances: dict[str, np.ndarray] = {}
    for _ in range(n_steps):
        x1, _y1 = next(it1)
        x2, _y2 = next(it2)
        if not x1.is_cuda:
            x1, x2 = x1.pin_memory().to(dev, non_blocking=True), x2.pin_memory().to(dev, non_blocking=True)
        update_distances(distances, est.pair_distances(x1, x2))
    if save_dir is not None:
        Path(save_dir).mkdir(parents=True, exist_ok=True)
        with open(Path(save_dir) / "distances.pkl", "wb") as f:
            pickle.dump(distances, f)
    return distances


def pair_distances(model, x1, x2) -> dict[str, np.ndarray]:
    return PlasticityEstimator(model).pair_distances(x1, x2)


def update_distances(distances: dict, new: dict) -> None:
    """Accumulate batches like the reference's ``update_dict`` (src/vitef/utils.py:208-213)."""
    for k, v in new.items():
        distances[k] = np.concatenate((distances[k], v), axis=0) if k in distances else v


def get_plasticity(distances: dict[str, np.ndarray]) -> dict[str, list[np.ndarray]]:
    """Ratio of every component's distance to the embedding distance, grouped per component in layer order
    (apps/plots/analysis.py:74-108; the reference reads ``distances.pkl``, this takes the dict directly)."""
    inputs = np.asarray(distances["embedding"]).flatten()
    out: dict[str, list[np.ndarray]] = {}
    for key, val in distances.items():
        if key == "embedding":
            continue
        _, comp = key.split("_", 1)
        out.setdefault(comp, []).append(np.asarray(val).flatten() / inputs)
    return out
