"""nn.Module tree of the ViT encoder, name-compatible with the reference and backed by the sm_100a kernels.

The module / attribute / ``state_dict`` names are API: the reference apps reach
into ``model.model.blocks[i].attn.qkv_mat`` etc. and ``freeze_model`` matches parameter names by substring
(apps/vit/utils.py:54-91). So parameters live in ordinary ``nn.Linear`` / ``nn.LayerNorm`` / ``nn.Conv2d`` holders with
the reference's names, while every ``forward`` routes through ``vit_plasticity_b200.ops`` (bf16 activations, fp32
accumulation, no CPU path). Only the ViT-relevant branches of the reference's generic Transformer are implemented
(hybrid image patching, cls token, learned positions, LayerNorm, pre-norm, GELU, classification head); other
settings raise ``NotImplementedError`` instead of silently running something else.

Reference: src/vitef/models/transformer/architecture.py and transformer/utils.py (line numbers cited per class).
"""

from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops
from .config import TransformerConfig


def _require_cuda(t: torch.Tensor, who: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{who}: input is on {t.device}; vit_plasticity_b200 runs on CUDA (sm_100a) only — there is no CPU fallback")


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm holder (transformer/utils.py:292-293) whose forward is the warp-per-row CUDA kernel."""

    def __init__(self, normalized_shape, eps: float = 1e-5, bias: bool = True):
        super().__init__(normalized_shape, eps=eps, bias=True)
        if not bias:
            raise NotImplementedError("LayerNorm without bias is not on the ViT path (vit.py:153 sets norm_bias=True)")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, "LayerNorm")
        return ops.LayerNormFn.apply(x, self.weight, self.bias, self.eps)


class Linear(nn.Linear):
    """nn.Linear holder whose forward is the tcgen05 GEMM with the bias in the epilogue."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, "Linear")
        return ops.LinearFn.apply(x, self.weight, self.bias, None)


class SelfAttention(nn.Module):
    """Multi-head self-attention (architecture.py:131-239): fused qkv Linear, attention core, output Linear."""

    def __init__(self, config: TransformerConfig):
        super().__init__()
        assert config.emb_dim % config.n_heads == 0, "Embedding dimension must be divisible by number of heads."
        if config.causal:
            raise NotImplementedError("causal attention is not on the ViT path (vit.py:148 sets causal=False)")
        if not config.attn_bias:
            raise NotImplementedError("attention without bias is not on the ViT path (vit.py:145 sets attn_bias=True)")
        if config.emb_dim // config.n_heads != 64:
            raise NotImplementedError("the fused attention kernel supports head_dim == 64 (ViT-B/L/H all use 64)")
        self.h = config.n_heads
        self.qkv_mat = Linear(config.emb_dim, 3 * config.emb_dim, bias=True)
        self.output = Linear(config.emb_dim, config.emb_dim, bias=True)
        self.flash = config.flash
        self.causal = False
        self.dropout = config.attn_dropout
        if self.dropout:
            raise NotImplementedError("attention dropout is 0 on the ViT path (vit.py:146)")

    def forward(self, x: torch.Tensor, verbose: bool = False, residual: torch.Tensor | None = None):
        _require_cuda(x, "SelfAttention")
        if verbose:
            return self._forward_with_maps(x)
        return ops.AttentionFn.apply(x, self.qkv_mat.weight, self.qkv_mat.bias, self.output.weight, self.output.bias, residual, self.h)

    def _forward_with_maps(self, x: torch.Tensor):
        """verbose=True (architecture.py:214-238) must return the (N,h,L,L) attention matrices, which the fused
        kernel never materialises. Debug-only slow path: projections on the GEMM kernel, the small L x L softmax in
        plain torch on device. Not used by train / analysis / probing."""
        n, l, e = x.shape
        d = e // self.h
        qkv = self.qkv_mat(x).float().view(n, l, 3, self.h, d).permute(2, 0, 3, 1, 4)
        attn = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) / math.sqrt(d), dim=-1)
        z = (attn @ qkv[2]).transpose(1, 2).reshape(n, l, e)
        return self.output(z.to(torch.bfloat16)), attn


class FeedForward(nn.Module):
    """fc1 -> exact-erf GELU -> fc2 (architecture.py:247-299); GELU lives in the fc1 GEMM epilogue."""

    def __init__(self, config: TransformerConfig):
        super().__init__()
        if config.activation.lower() != "gelu":
            raise NotImplementedError(f"activation '{config.activation}' is not on the ViT path (vit.py:149 sets gelu)")
        if not config.ffn_bias:
            raise NotImplementedError("feed-forward without bias is not on the ViT path (vit.py:150)")
        self.fc1 = Linear(config.emb_dim, config.ffn_dim, bias=True)
        self.fc2 = Linear(config.ffn_dim, config.emb_dim, bias=True)
        self.activation = torch.nn.functional.gelu  # kept for API parity; the kernels apply it in the epilogue
        self.dropout = config.ffn_dropout
        if self.dropout:
            raise NotImplementedError("feed-forward dropout is 0 on the ViT path (vit.py:151)")

    def forward(self, x: torch.Tensor, residual: torch.Tensor | None = None) -> torch.Tensor:
        _require_cuda(x, "FeedForward")
        return ops.MlpFn.apply(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, residual)


class TransformerBlock(nn.Module):
    """Pre-norm transformer block (architecture.py:307-383) plus its decomposition / probing taps (385-502)."""

    def __init__(self, config: TransformerConfig):
        super().__init__()
        if config.norm.lower() != "layer":
            raise NotImplementedError(f"norm '{config.norm}' is not on the ViT path (vit.py:152 sets layer)")
        if not config.pre_norm:
            raise NotImplementedError("post-norm blocks are not on the ViT path (vit.py:155 sets pre_norm=True)")
        self.attn_norm = LayerNorm(config.emb_dim, eps=config.norm_eps, bias=config.norm_bias)
        self.attn = SelfAttention(config)
        self.ffn_norm = LayerNorm(config.emb_dim, eps=config.norm_eps, bias=config.norm_bias)
        self.ffn = FeedForward(config)
        self.pre_norm = True

    def forward(self, x: torch.Tensor, verbose: bool = False):
        _require_cuda(x, "TransformerBlock")
        if verbose:
            out, att = self.attn(self.attn_norm(x), verbose=True)
            out = x.to(out.dtype) + out
            return self.ffn(self.ffn_norm(out), residual=out), att
        a, f = self.attn, self.ffn
        return ops.BlockFn.apply(
            x, self.attn_norm.weight, self.attn_norm.bias, a.qkv_mat.weight, a.qkv_mat.bias, a.output.weight, a.output.bias,
            self.ffn_norm.weight, self.ffn_norm.bias, f.fc1.weight, f.fc1.bias, f.fc2.weight, f.fc2.bias, a.h, self.attn_norm.eps,
        )

    @torch.inference_mode()
    def _decompose(self, x: torch.Tensor) -> dict:
        """Five independent maps of the same input, returned as fp32 CPU tensors (architecture.py:385-418).
        fc2 on [x,0,0,0] equals x @ W2[:, :E]^T + b2, so the zero-padded FLOPs are skipped."""
        e = x.shape[-1]
        outputs = {}
        outputs["attn_norm"] = self.attn_norm(x).float().cpu()
        outputs["attn"] = self.attn(x).float().cpu()
        outputs["ffn_norm"] = self.ffn_norm(x).float().cpu()
        outputs["ffn_fc1"] = self.ffn.fc1(x).float().cpu()
        w2 = ops.shadow_bf16(self.ffn.fc2.weight)[:, :e]
        x2 = ops._as_bf16_2d(x)
        y = torch.empty(x2.shape[0], e, device=x.device, dtype=torch.bfloat16)
        L.gemm(x2, w2, m=x2.shape[0], n=e, k=e, epilogue=L.EPI_BF16, bias=self.ffn.fc2.bias.detach(), out=y)
        outputs["ffn_fc2"] = y.view(*x.shape[:-1], e).float().cpu()
        return outputs

    @torch.inference_mode()
    def _probes(self, x: torch.Tensor):
        """Chained forward with the reference's 8 taps (architecture.py:436-467), fp32 CPU tensors."""
        probes = {}
        out = self.attn_norm(x)
        probes["attn_norm"] = out.float().cpu()
        out = self.attn(out)
        probes["attn"] = out.float().cpu()
        out_res = L.add_bf16(ops._as_bf16_2d(x), ops._as_bf16_2d(out)).view(out.shape)
        probes["attn_res"] = out_res.float().cpu()
        out = self.ffn_norm(out_res)
        probes["ffn_norm"] = out.float().cpu()
        h2 = ops._as_bf16_2d(out)
        act, z = ops.linear_fwd(h2, ops.shadow_bf16(self.ffn.fc1.weight), self.ffn.fc1.bias.detach(), gelu=True)
        probes["ffn_fc1"] = z.view(*out.shape[:-1], -1).float().cpu()
        probes["ffn_activation"] = act.view(*out.shape[:-1], -1).float().cpu()
        out = self.ffn.fc2(act).view(out_res.shape)
        probes["ffn_fc2"] = out.float().cpu()
        out = L.add_bf16(ops._as_bf16_2d(out_res), ops._as_bf16_2d(out)).view(out_res.shape)
        probes["ffn_res"] = out.float().cpu()
        return out, probes


    @torch.inference_mode()
    def _pooled_probes(self, x: torch.Tensor, cls_pooling: bool, normalize: bool):
        """The 8 taps of ``_probes`` pooled on the device (cls row or token mean, L2-normalised): what
        apps/vit/linear_probing.py:92-112 computes on the host after copying every (N, L, D) tap over PCIe."""
        n, seq, _ = x.shape
        pool = lambda t: L.pool_tokens(ops._as_bf16_2d(t), n, seq, cls_pooling, normalize)
        probes = {}
        out = self.attn_norm(x)
        probes["attn_norm"] = pool(out)
        out = self.attn(out)
        probes["attn"] = pool(out)
        out_res = L.add_bf16(ops._as_bf16_2d(x), ops._as_bf16_2d(out)).view(out.shape)
        probes["attn_res"] = pool(out_res)
        out = self.ffn_norm(out_res)
        probes["ffn_norm"] = pool(out)
        act, z = ops.linear_fwd(ops._as_bf16_2d(out), ops.shadow_bf16(self.ffn.fc1.weight), self.ffn.fc1.bias.detach(), gelu=True)
        probes["ffn_fc1"] = pool(z)
        probes["ffn_activation"] = pool(act)
        out = self.ffn.fc2(act).view(out_res.shape)
        probes["ffn_fc2"] = pool(out)
        out = L.add_bf16(ops._as_bf16_2d(out_res), ops._as_bf16_2d(out)).view(out_res.shape)
        probes["ffn_res"] = pool(out)
        return out, probes


class PatchImages(nn.Module):
    """Hybrid patching (transformer/utils.py:38-115): a Conv2d with kernel = stride = P holds the weights
    (``patching.0``, shape (E, C, P, P)); the forward is im2col + the tcgen05 GEMM."""

    def __init__(self, image_dim: tuple, image_patch: str, patch_size: int, emb_dim: int):
        super().__init__()
        n_channels, height, width = image_dim
        assert (height % patch_size == 0) and (width % patch_size == 0), "Image dimensions must be divisible by the patch size."
        if image_patch.lower() != "hybrid":
            raise NotImplementedError("only image_patch='hybrid' is on the ViT path (vit.py:140)")
        if patch_size % 8 != 0:
            raise NotImplementedError("the im2col kernel needs patch_size % 8 == 0 (ViT-B/L use 16)")
        self.n_patches = height * width // (patch_size**2)
        self.patch_dim = patch_size**2 * n_channels
        self.patch_size = patch_size
        self.patching = nn.Sequential(
            nn.Conv2d(in_channels=n_channels, out_channels=emb_dim, kernel_size=patch_size, stride=patch_size),
            nn.Flatten(start_dim=2),
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, "PatchImages")
        conv = self.patching[0]
        patches = L.im2col_patches(x.float().contiguous(), self.patch_size)
        out = ops.LinearFn.apply(patches, conv.weight.view(conv.weight.shape[0], -1), conv.bias, None)
        return out.view(x.shape[0], self.n_patches, -1)


class Embedding(nn.Module):
    """Patch embedding + cls token + learned positions (architecture.py:510-678)."""

    def __init__(self, config: TransformerConfig):
        super().__init__()
        if (config.patch_type or "").lower() != "computer_vision":
            raise NotImplementedError("only patch_type='computer_vision' is on the ViT path (vit.py:139)")
        self.patching = PatchImages(image_dim=config.image_dim, image_patch=config.image_patch, patch_size=config.patch_size, emb_dim=config.emb_dim)
        config.seq_len = self.patching.n_patches
        config.vocab_size = self.patching.patch_dim
        if not config.cls_token or not config.pos_emb:
            raise NotImplementedError("the ViT path uses a cls token and positional embeddings (vit.py:142,156)")
        self.cls_token = nn.Parameter(torch.randn(1, 1, config.emb_dim))
        config.seq_len += 1
        self.token_emb = nn.Identity()
        self.L = config.seq_len
        self.pos_dim = config.emb_dim
        self.pos_emb = nn.Parameter(torch.randn(1, self.L, self.pos_dim)).requires_grad_(not config.freeze_pos)
        self.dropout = config.emb_dropout
        if self.dropout:
            raise NotImplementedError("embedding dropout is 0 on the ViT path (vit.py:144)")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, "Embedding")
        conv = self.patching.patching[0]
        if x.dim() == 2 and x.dtype == torch.bfloat16:
            # bf16 patch rows [N * n_patches, C*P*P] straight from the device-side input pipeline
            # (preprocess.DevicePreprocessor.patches): the im2col pass and the fp32 image are skipped
            return ops.EmbedFn.apply(x, conv.weight, conv.bias, self.cls_token, self.pos_emb, self.patching.patch_size)
        if x.dtype != torch.float32:
            x = x.float()
        return ops.EmbedFn.apply(x, conv.weight, conv.bias, self.cls_token, self.pos_emb, self.patching.patch_size)


class ClassificationLayer(nn.Module):
    """LayerNorm then Linear on the cls token (transformer/utils.py:355-422). The reference normalises every token
    and then reads row 0; only row 0 is normalised here (same result, 1/L of the work)."""

    def __init__(self, emb_dim: int, n_classes: int, norm: str, norm_eps: float, norm_bias: bool, dropout: float):
        super().__init__()
        if norm.lower() != "layer":
            raise NotImplementedError(f"norm '{norm}' is not on the ViT path")
        self.output_norm = LayerNorm(emb_dim, eps=norm_eps, bias=norm_bias)
        self.dropout = dropout
        if self.dropout:
            raise NotImplementedError("output dropout is 0 on the ViT path (vit.py:160)")
        self.output = nn.Linear(emb_dim, n_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, "ClassificationLayer")
        cls_rows = self.output_norm(x[:, 0, :].contiguous())
        # (N x E) @ (E x n_classes) with n_classes ~ 10: 0.00002 % of the step's FLOPs, kept in fp32
        return torch.nn.functional.linear(cls_rows.float(), self.output.weight, self.output.bias)


class Output(nn.Module):
    """Task head container (architecture.py:686-775); only the classification head is on the ViT path."""

    def __init__(self, config: TransformerConfig):
        super().__init__()
        self.output_type = config.output_type
        if self.output_type.lower() != "classification":
            raise NotImplementedError(f"output_type '{config.output_type}' is not on the ViT path (vit.py:158)")
        self.output_layer = ClassificationLayer(
            emb_dim=config.emb_dim, n_classes=config.n_classes, norm=config.norm, norm_eps=config.norm_eps,
            norm_bias=config.norm_bias, dropout=config.output_dropout,
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.output_layer(x)


class Transformer(nn.Module):
    """Embedding -> n_layers x TransformerBlock -> Output (architecture.py:783-911)."""

    def __init__(self, config: TransformerConfig):
        super().__init__()
        self.embedding = Embedding(config)
        self.blocks = nn.ModuleList([TransformerBlock(config) for _ in range(config.n_layers)])
        self.output = Output(config)

    def forward(self, x: torch.Tensor, verbose: bool = False):
        out = self.embedding(x)
        attentions = []
        for block in self.blocks:
            out = block(out, verbose=verbose)
            if verbose:
                out, att = out
                attentions.append(att)
        out = self.output(out)
        if verbose:
            return out, torch.stack(attentions)
        return out

    @torch.inference_mode()
    def get_decomposition(self, x: torch.Tensor) -> dict:
        """Drop-in for architecture.py:856-883: dict of fp32 CPU tensors, every block fed the SAME embedding
        output. This API is PCIe-bound by construction; the fused on-device estimator is
        ``vit_plasticity_b200.plasticity.pair_distances``."""
        outputs = {}
        out = self.embedding(x)
        outputs["embedding"] = out.float().cpu()
        for i, block in enumerate(self.blocks):
            for key, val in block._decompose(out).items():
                outputs[f"block{i}_{key}"] = val
        return outputs

    @torch.inference_mode()
    def get_pooled_probes(self, x: torch.Tensor, cls_pooling: bool = True, normalize: bool = True) -> dict:
        """``get_probes`` followed by the pooling + normalisation of linear_probing.get_embeddings, done on the device:
        {key: (N, D) float32 CPU tensor}. 96 x (N, D) rows cross PCIe per batch instead of 96 x (N, 197, D) tensors."""
        probes = {}
        out = self.embedding(x)
        for i, block in enumerate(self.blocks):
            out, block_probes = block._pooled_probes(out, cls_pooling, normalize)
            for key, val in block_probes.items():
                probes[f"block{i}_{key}"] = val
        return {k: v.cpu() for k, v in probes.items()}

    @torch.inference_mode()
    def get_probes(self, x: torch.Tensor) -> dict:
        """Drop-in for architecture.py:885-911 (chained, 8 taps per block)."""
        probes = {}
        out = self.embedding(x)
        for i, block in enumerate(self.blocks):
            out, block_probes = block._probes(out)
            for key, val in block_probes.items():
                probes[f"block{i}_{key}"] = val
        return probes
