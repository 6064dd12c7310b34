"""Fused, on-device per-component plasticity estimator.

Reference path (what this replaces): ``Transformer.get_decomposition`` twice (architecture.py:856-883, every
component output copied to the host), then per key an H2D re-upload and ``distance`` (apps/vit/analysis.py:216-233,
:68), then offline ``ratio = dist[key] / dist["embedding"]`` (apps/plots/analysis.py:97).

Here nothing but the final ``(1 + 5 * n_layers) x N`` table of distances leaves the GPU and no component output or
difference tensor is ever written to HBM:

* every block sees the SAME embedding output e (architecture.py:877-881), so the per-layer weights are concatenated
  and each component family is ONE GEMM over all layers;
* linear components — patch embedding, ``fc1``, ``fc2`` on ``[e,0,0,0]`` (== ``W2[:, :E]``) and the attention output
  projection — satisfy f(a) - f(b) = W (a - b) (the bias cancels), so the GEMM runs on the difference, computed in
  fp32 BEFORE the bf16 down-cast, and its epilogue reduces sum(acc^2) per sample straight out of TMEM
  (``VB_EPI_SUMSQ``);
* LayerNorm: LN_i(a) - LN_i(b) = gamma_i * (zhat_a - zhat_b), so one fp32 kernel produces
  u[s, d] = sum_l (zhat_a - zhat_b)^2 and every norm's squared distance is u @ gamma_i^2;
* attention is non-linear: q/k/v for both inputs come from one concatenated-weight GEMM each, the paired attention
  kernel subtracts the two head outputs in fp32, and the (linear) output projection runs on that difference with the
  sum-of-squares epilogue.
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
    @torch.no_grad()
    def squared_distances(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        """f32 device tensor [1 + 5 * n_layers, N] of SQUARED distances, rows ordered like the reference's keys."""
        if not (x1.is_cuda and x2.is_cuda):
            raise RuntimeError("PlasticityEstimator runs on CUDA only — there is no CPU fallback")
        self._refresh()
        x1, x2 = x1.float().contiguous(), x2.float().contiguous()
        n = x1.shape[0]
        e, f, nl, heads = self.e, self.f, self.n_layers, self.heads
        dev = x1.device
        # ---- embedding of both inputs (fp32 out of the accumulator) and of the difference ----
        p1, p2 = L.im2col_patches(x1, self.patch), L.im2col_patches(x2, self.patch)
        pd = L.im2col_patches(x1, self.patch, x2)  # (x1 - x2) subtracted in fp32, then bf16
        rows_p, kp = p1.shape
        np_ = rows_p // n
        seq = np_ + 1
        m = n * seq
        emb = []
        for p in (p1, p2):
            po = torch.empty(rows_p, e, device=dev, dtype=torch.float32)
            L.gemm(p, self.w_patch, m=rows_p, n=e, k=kp, epilogue=L.EPI_F32, bias=self.b_patch, out=po)
            emb.append(L.assemble_tokens(None, po, self.cls, self.pos, n, np_, e, want_bf16=True, want_f32=True))
        (t1_16, t1_32), (t2_16, t2_32) = emb
        ss_emb = torch.zeros(n, 1, device=dev, dtype=torch.float32)
        L.gemm(pd, self.w_patch, m=rows_p, n=e, k=kp, epilogue=L.EPI_SUMSQ, sumsq=ss_emb, rows_per_sample=np_, cols_per_group=e, n_groups=1)
        dpo = torch.empty(rows_p, e, device=dev, dtype=torch.bfloat16)
        L.gemm(pd, self.w_patch, m=rows_p, n=e, k=kp, epilogue=L.EPI_BF16, out=dpo)
        dtok, _ = L.assemble_tokens(dpo, None, self.zeros_cls, self.zeros_pos, n, np_, e)  # cls row of the difference is 0
        # ---- LayerNorms: all 2 * n_layers at once ----
        u = torch.zeros(n, e, device=dev, dtype=torch.float32)
        L.layernorm_pair_sqdiff(t1_32, t2_32, u, n, seq, e, self.eps)
        ss_ln = u @ self.gamma_sq  # [N, 2 * n_layers]; 2*N*E*2n_layers FLOPs, negligible
        # ---- fc1 / fc2 on the difference, all layers in one GEMM each ----
        ss_fc1 = torch.zeros(n, nl, device=dev, dtype=torch.float32)
        L.gemm(dtok, self.w_fc1, m=m, n=nl * f, k=e, epilogue=L.EPI_SUMSQ, sumsq=ss_fc1, rows_per_sample=seq, cols_per_group=f, n_groups=nl)
        ss_fc2 = torch.zeros(n, nl, device=dev, dtype=torch.float32)
        L.gemm(dtok, self.w_fc2, m=m, n=nl * e, k=e, epilogue=L.EPI_SUMSQ, sumsq=ss_fc2, rows_per_sample=seq, cols_per_group=e, n_groups=nl)
        # ---- attention (applied to the un-normalised embedding, architecture.py:405) ----
        ss_attn = torch.zeros(n, nl, device=dev, dtype=torch.float32)
        qkv = []
        for t in (t1_16, t2_16):
            q = torch.empty(m, nl * 3 * e, device=dev, dtype=torch.bfloat16)
            L.gemm(t, self.w_qkv, m=m, n=nl * 3 * e, k=e, epilogue=L.EPI_BF16, bias=self.b_qkv, out=q)
            qkv.append(q)
        if seq <= 208:
            delta_all = L.attention_pair_delta_layers(qkv[0], qkv[1], nl, n, seq, heads, e // heads)  # one launch, all layers
        else:
            delta_all = torch.empty(nl, m, e, device=dev, dtype=torch.bfloat16)
            for i in range(nl):
                qa, qb = qkv[0][:, i * 3 * e : (i + 1) * 3 * e], qkv[1][:, i * 3 * e : (i + 1) * 3 * e]
                L.attention_pair_delta(qa, qb, delta_all[i], n, seq, heads, e // heads)
        for i in range(nl):
            # sumsq pointer offset by i with n_groups = n_layers writes column i of ss_attn
            L.gemm(delta_all[i], self.w_out[i], m=m, n=e, k=e, epilogue=L.EPI_SUMSQ, sumsq=ss_attn[:, i:], rows_per_sample=seq, cols_per_group=e, n_groups=nl)
        # ---- assemble in the reference's key order ----
        rows = [ss_emb[:, 0]]
        for i in range(nl):
            rows += [ss_ln[:, i], ss_attn[:, i], ss_ln[:, nl + i], ss_fc1[:, i], ss_fc2[:, i]]
        return torch.stack(rows, 0)

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


def gather_tables(local: np.ndarray, n_total: int, group=None) -> np.ndarray | None:
    """Concatenate the ranks' [rows, n_local] distance tables along the pair axis on rank 0 (contiguous shards in rank
    order, see ``distributed.shard_range``). The only collective of the sweep: one gather at the end."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0, group=group)
    if rank != 0:
        return None
    table = np.concatenate(parts, axis=1)
    assert table.shape[1] == n_total, (table.shape, n_total)
    return table


def perturbation_sweep(model, images, eps_list, noise_seed: int = 0, pairs_per_call: int = 64, rank: int | None = None,
                       world: int | None = None, group=None, estimator: "PlasticityEstimator | None" = None):
    """Plasticity under input perturbations of growing magnitude (BASELINE.json configs[4]): pairs (x, x + eps * n),
    n ~ N(0, 1) drawn per image from ``noise_seed``, for every eps in ``eps_list``.

    The N images are sharded contiguously over the ranks (each pair is independent: apps/vit/analysis.py:68 reduces per
    sample); every rank streams its shard from ``images`` (host or device, fp32 NCHW) through the fused estimator and
    only the (1 + 5 n_layers) x N_local distance table per eps leaves the GPU. No collective on the data path; rank 0
    receives {eps: {key: (N,) float32}} from one gather per eps at the end, the other ranks get None.
    """
    from .distributed import shard_range

    est = estimator if estimator is not None else PlasticityEstimator(model)
    n_total = images.shape[0]
    lo, hi = shard_range(n_total, rank, world)
    dev = next(_inner(model).parameters()).device
    tables = {float(e): [] for e in eps_list}
    for s0 in range(lo, hi, pairs_per_call):
        s1 = min(s0 + pairs_per_call, hi)
        x = images[s0:s1].to(dev, non_blocking=True).float()
        # per-image generators: the noise of image i does not depend on how the images are sharded or batched
        noise = torch.stack([torch.randn(x.shape[1:], generator=torch.Generator().manual_seed(noise_seed * 1_000_003 + i)) for i in range(s0, s1)]).to(dev)
        for e in eps_list:
            tables[float(e)].append(est.squared_distances(x, x + float(e) * noise).sqrt())
    out = {}
    keys = est.keys()
    for e, chunks in tables.items():
        local = torch.cat(chunks, 1).cpu().numpy() if chunks else np.zeros((len(keys), 0), np.float32)
        table = gather_tables(local, n_total, group)
        out[e] = None if table is None else {k: table[j] for j, k in enumerate(keys)}
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_rank(group) != 0:
        return None
    return out


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
