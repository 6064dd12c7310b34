"""CPU oracle for the ViT hot path of ambroiseodt/vit-plasticity — TEST INFRASTRUCTURE ONLY.

A plain PyTorch fp32, functional restatement (no nn.Module tree, operates on a ``state_dict``) of the reference
algorithm for: ViT forward, cross-entropy finetuning step (forward + backward + clip + SGD-momentum), the
per-component decomposition, the per-sample Frobenius ``distance`` and the plasticity ratio. Each function cites
the reference file:line it follows (paths relative to /root/reference).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker / CPU baseline — never on the product path.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build container: ``oracle/make_golden.py``
imports ``vitef`` from /root/reference, loads identical weights into the reference modules and into this oracle, and
(a) asserts they agree to fp32 round-off, (b) writes the reference's outputs to ``tests/golden/*.pt``. The CPU test
suite then re-checks the oracle against those committed fixtures.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

# src/vitef/models/vit.py:130-134
VIT_SIZES = {
    "base": dict(emb_dim=768, n_heads=12, n_layers=12, ffn_dim=3072),
    "large": dict(emb_dim=1024, n_heads=16, n_layers=24, ffn_dim=4096),
    "huge": dict(emb_dim=1280, n_heads=16, n_layers=32, ffn_dim=5120),
}
LN_EPS = 1e-12  # src/vitef/models/vit.py:154


@dataclass
class Arch:
    emb_dim: int
    n_heads: int
    n_layers: int
    ffn_dim: int
    patch_size: int = 16
    image_dim: tuple = (3, 224, 224)
    n_classes: int = 10
    norm_eps: float = LN_EPS

    @property
    def n_patches(self) -> int:  # transformer/utils.py:83
        return self.image_dim[1] * self.image_dim[2] // self.patch_size**2

    @property
    def seq_len(self) -> int:  # architecture.py:593,603 (+1 for the cls token)
        return self.n_patches + 1


def vit_arch(model_name: str = "base", n_classes: int = 10, patch_size: int = 16, image_dim=(3, 224, 224)) -> Arch:
    return Arch(**VIT_SIZES[model_name], patch_size=patch_size, image_dim=tuple(image_dim), n_classes=n_classes)


# --------------------------------------------------------------------------------------------------
# Deterministic random init with the reference's parameter names/shapes.
# Distributions follow torch's defaults for the layers the reference instantiates (nn.Linear / nn.Conv2d:
# U(+-1/sqrt(fan_in)); nn.LayerNorm: ones/zeros; cls_token / pos_emb: randn, architecture.py:602,635).
# The exact draw order is NOT the reference's; parity tests always copy one state_dict into both sides.
# --------------------------------------------------------------------------------------------------
def init_state_dict(arch: Arch, seed: int = 42, prefix: str = "") -> dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    E, F_, P, C = arch.emb_dim, arch.ffn_dim, arch.patch_size, arch.image_dim[0]

    def uniform(shape, fan_in):
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    sd: dict[str, torch.Tensor] = {}
    sd["embedding.cls_token"] = torch.randn(1, 1, E, generator=g)
    sd["embedding.pos_emb"] = torch.randn(1, arch.seq_len, E, generator=g)
    sd["embedding.patching.patching.0.weight"] = uniform((E, C, P, P), C * P * P)
    sd["embedding.patching.patching.0.bias"] = uniform((E,), C * P * P)
    for i in range(arch.n_layers):
        b = f"blocks.{i}."
        for norm in ("attn_norm", "ffn_norm"):
            # non-trivial affine so that tests exercise gamma/beta (a fresh nn.LayerNorm would be ones/zeros)
            sd[b + norm + ".weight"] = 1.0 + 0.1 * torch.randn(E, generator=g)
            sd[b + norm + ".bias"] = 0.1 * torch.randn(E, generator=g)
        sd[b + "attn.qkv_mat.weight"] = uniform((3 * E, E), E)
        sd[b + "attn.qkv_mat.bias"] = uniform((3 * E,), E)
        sd[b + "attn.output.weight"] = uniform((E, E), E)
        sd[b + "attn.output.bias"] = uniform((E,), E)
        sd[b + "ffn.fc1.weight"] = uniform((F_, E), E)
        sd[b + "ffn.fc1.bias"] = uniform((F_,), E)
        sd[b + "ffn.fc2.weight"] = uniform((E, F_), F_)
        sd[b + "ffn.fc2.bias"] = uniform((E,), F_)
    sd["output.output_layer.output_norm.weight"] = 1.0 + 0.1 * torch.randn(E, generator=g)
    sd["output.output_layer.output_norm.bias"] = 0.1 * torch.randn(E, generator=g)
    sd["output.output_layer.output.weight"] = uniform((arch.n_classes, E), E)
    sd["output.output_layer.output.bias"] = uniform((arch.n_classes,), E)
    return {prefix + k: v for k, v in sd.items()}


def synthetic_images(n: int, arch: Arch, seed: int) -> torch.Tensor:
    """CIFAR-10-shaped inputs after the reference transform (resize to 224 + ImageNet normalisation,
    src/vitef/data/images/utils.py:337-366) are ~zero-mean/unit-variance float32 NCHW: modelled as randn."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(n, *arch.image_dim, generator=g)


def synthetic_labels(n: int, arch: Arch, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randint(0, arch.n_classes, (n,), generator=g)


# --------------------------------------------------------------------------------------------------
# Forward pieces
# --------------------------------------------------------------------------------------------------
def embedding(sd, x: torch.Tensor, arch: Arch) -> torch.Tensor:
    """Embedding.forward, architecture.py:644-678: Conv2d(k=s=P) -> flatten -> transpose
    (transformer/utils.py:91,114), prepend cls (666-669), add pos_emb (672-675); dropout p=0."""
    P = arch.patch_size
    w, b = sd["embedding.patching.patching.0.weight"], sd["embedding.patching.patching.0.bias"]
    tok = F.conv2d(x, w, b, stride=P).flatten(2).transpose(1, 2)  # (N, n_patches, E)
    cls = sd["embedding.cls_token"].expand(x.shape[0], -1, -1)
    return torch.cat((cls, tok), dim=1) + sd["embedding.pos_emb"]


def layer_norm(sd, key: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    """nn.LayerNorm over the last dim with affine (transformer/utils.py:293; eps from vit.py:154)."""
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], eps)


def attention(sd, key: str, x: torch.Tensor, n_heads: int) -> torch.Tensor:
    """SelfAttention.forward, architecture.py:189-239 (vanilla branch: flash=False, causal=False, dropout 0)."""
    N, L, E = x.shape
    d = E // n_heads
    qkv = F.linear(x, sd[key + ".qkv_mat.weight"], sd[key + ".qkv_mat.bias"])  # :205
    q, k, v = (t.view(N, L, n_heads, d).transpose(1, 2) for t in qkv.chunk(3, dim=-1))  # :205-212
    scores = q @ k.transpose(-1, -2) / math.sqrt(d)  # :217
    z = torch.softmax(scores, dim=-1) @ v  # :223-226
    z = z.transpose(1, 2).reshape(N, L, E)  # :233
    return F.linear(z, sd[key + ".output.weight"], sd[key + ".output.bias"])  # :236


def feed_forward(sd, key: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward.forward, architecture.py:295-298: fc1 -> exact-erf GELU (273-274) -> fc2."""
    h = F.gelu(F.linear(x, sd[key + ".fc1.weight"], sd[key + ".fc1.bias"]))
    return F.linear(h, sd[key + ".fc2.weight"], sd[key + ".fc2.bias"])


def block(sd, i: int, x: torch.Tensor, arch: Arch) -> torch.Tensor:
    """TransformerBlock.forward pre-norm branch, architecture.py:369-374."""
    b = f"blocks.{i}."
    out = x + attention(sd, b + "attn", layer_norm(sd, b + "attn_norm", x, arch.norm_eps), arch.n_heads)
    return out + feed_forward(sd, b + "ffn", layer_norm(sd, b + "ffn_norm", out, arch.norm_eps))


def forward(sd, x: torch.Tensor, arch: Arch) -> torch.Tensor:
    """Transformer.forward, architecture.py:824-854, with ClassificationLayer.forward
    (transformer/utils.py:416-420: LayerNorm over all tokens, Linear on token 0)."""
    out = embedding(sd, x, arch)
    for i in range(arch.n_layers):
        out = block(sd, i, out, arch)
    out = layer_norm(sd, "output.output_layer.output_norm", out, arch.norm_eps)
    return F.linear(out[:, 0, :], sd["output.output_layer.output.weight"], sd["output.output_layer.output.bias"])


# --------------------------------------------------------------------------------------------------
# Finetuning step (apps/vit/train.py:263-283) with component freezing (apps/vit/utils.py:54-91)
# --------------------------------------------------------------------------------------------------
FREEZE_MAP = {  # apps/vit/utils.py:67-74
    "emb": ["embedding"],
    "attn_norm": ["attn_norm"],
    "mha": ["attn.qkv_mat", "attn.output"],
    "ffn_norm": ["ffn_norm"],
    "ffn_fc1": ["ffn.fc1"],
    "ffn_fc2": ["ffn.fc2"],
}


def frozen_keys(sd, components) -> set[str]:
    """Names frozen by freeze_model: 'emb' freezes every embedding parameter (:82-84); block parameters are
    matched by substring (:87-91); the final norm and head are never frozen."""
    pats = [p for c in components for p in FREEZE_MAP[c]]
    out = set()
    for k in sd:
        if k.startswith("embedding."):
            if "embedding" in pats:
                out.add(k)
        elif k.startswith("blocks.") and any(p in k for p in pats):
            out.add(k)
    return out


def loss_and_grads(sd, x, y, arch: Arch, frozen: set[str] = frozenset()):
    """Cross-entropy loss (train.py:264) and d loss / d param for every non-frozen parameter (train.py:270)."""
    params = {k: v.detach().clone().requires_grad_(k not in frozen) for k, v in sd.items()}
    logits = forward(params, x, arch)
    loss = F.cross_entropy(logits, y)
    leaves = [k for k in params if params[k].requires_grad]
    grads = torch.autograd.grad(loss, [params[k] for k in leaves])
    return loss.detach(), logits.detach(), dict(zip(leaves, grads))


def sgd_step(sd, bufs, grads, lr: float, momentum: float, grad_clip: float | None):
    """clip_grad_norm_ (train.py:277-278; max-norm inf when grad_clip is None) then torch.optim.SGD with momentum
    (optim.py:76-82; dampening 0, no weight decay, no nesterov). Updates ``sd`` and ``bufs`` in place; returns the
    pre-clip total gradient norm (what the reference logs as grad_norm)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    max_norm = float("inf") if grad_clip is None else float(grad_clip)
    coef = min(1.0, max_norm / (float(total) + 1e-6))  # torch.nn.utils.clip_grad_norm_
    for k, g in grads.items():
        g = g * coef
        if momentum:
            if k not in bufs:
                bufs[k] = g.clone()  # first step: buf = grad
            else:
                bufs[k].mul_(momentum).add_(g)
            g = bufs[k]
        sd[k] = sd[k] - lr * g
    return total


# --------------------------------------------------------------------------------------------------
# Plasticity estimator
# --------------------------------------------------------------------------------------------------
def decomposition(sd, x: torch.Tensor, arch: Arch) -> dict[str, torch.Tensor]:
    """Transformer.get_decomposition (architecture.py:856-883) + TransformerBlock._decompose (385-418).
    Every block receives the SAME embedding output (877-881); attention is applied to the un-normalised input
    (405); fc2 sees [x, 0, 0, 0] (414-416), which requires ffn_dim == 4 * emb_dim."""
    out = {}
    e = embedding(sd, x, arch)
    out["embedding"] = e
    for i in range(arch.n_layers):
        b = f"blocks.{i}."
        out[f"block{i}_attn_norm"] = layer_norm(sd, b + "attn_norm", e, arch.norm_eps)
        out[f"block{i}_attn"] = attention(sd, b + "attn", e, arch.n_heads)
        out[f"block{i}_ffn_norm"] = layer_norm(sd, b + "ffn_norm", e, arch.norm_eps)
        out[f"block{i}_ffn_fc1"] = F.linear(e, sd[b + "ffn.fc1.weight"], sd[b + "ffn.fc1.bias"])
        zero = torch.zeros_like(e)
        expanded = torch.cat((e, zero, zero, zero), dim=-1)
        out[f"block{i}_ffn_fc2"] = F.linear(expanded, sd[b + "ffn.fc2.weight"], sd[b + "ffn.fc2.bias"])
    return out


def distance(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """apps/vit/analysis.py:68 with reduction='none': per-sample Frobenius norm over tokens and features."""
    return ((a - b) ** 2).flatten(start_dim=1).sum(dim=-1).sqrt()


def pair_distances(sd, x1, x2, arch: Arch) -> dict[str, np.ndarray]:
    """One iteration of the analysis loop, apps/vit/analysis.py:216-233 (both decompositions, distance per key)."""
    with torch.no_grad():
        o1, o2 = decomposition(sd, x1, arch), decomposition(sd, x2, arch)
        return {k: distance(o1[k], o2[k]).numpy() for k in o1}


def plasticity(distances: dict[str, np.ndarray]) -> dict[str, list[np.ndarray]]:
    """get_plasticity, apps/plots/analysis.py:74-108: ratio to the 'embedding' distance (:97), grouped per
    component in layer order."""
    inputs = np.asarray(distances["embedding"]).flatten()
    out: dict[str, list[np.ndarray]] = {}
    for key, val in distances.items():
        if key == "embedding":
            continue
        _, comp = key.split("_", 1)
        out.setdefault(comp, []).append(np.asarray(val).flatten() / inputs)
    return out


def probes(sd, x: torch.Tensor, arch: Arch) -> dict[str, torch.Tensor]:
    """Transformer.get_probes (architecture.py:885-911) + TransformerBlock._probes pre-norm branch (436-467):
    chained block forward with 8 taps per block."""
    taps = {}
    out = embedding(sd, x, arch)
    for i in range(arch.n_layers):
        b = f"blocks.{i}."
        h = layer_norm(sd, b + "attn_norm", out, arch.norm_eps)
        taps[f"block{i}_attn_norm"] = h
        h = attention(sd, b + "attn", h, arch.n_heads)
        taps[f"block{i}_attn"] = h
        res = out + h
        taps[f"block{i}_attn_res"] = res
        h = layer_norm(sd, b + "ffn_norm", res, arch.norm_eps)
        taps[f"block{i}_ffn_norm"] = h
        h = F.linear(h, sd[b + "ffn.fc1.weight"], sd[b + "ffn.fc1.bias"])
        taps[f"block{i}_ffn_fc1"] = h
        h = F.gelu(h)
        taps[f"block{i}_ffn_activation"] = h
        h = F.linear(h, sd[b + "ffn.fc2.weight"], sd[b + "ffn.fc2.bias"])
        taps[f"block{i}_ffn_fc2"] = h
        out = res + h
        taps[f"block{i}_ffn_res"] = out
    return taps
