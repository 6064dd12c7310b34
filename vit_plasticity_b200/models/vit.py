"""ViT wrapper, interface-compatible with src/vitef/models/vit.py:83-316.

Keeps: the size table (vit.py:130-134), the fixed architecture flags (136-162), ``.model`` / ``.config`` /
``.model_name`` attributes, ``forward`` / ``get_decomposition`` / ``get_probes`` bound to the inner Transformer as
instance attributes (173-177), the 2-class head when ``in21k`` (161) and the fresh finetuning head (235-237).
Loading pretrained weights from a local ``save_dir/<model_name>.pt`` works as in the reference (214-225); the
HuggingFace download branch (239-303) is out of scope (no network; BASELINE configs are random-init): a missing local
file raises FileNotFoundError instead of silently keeping the random initialisation.
"""

from __future__ import annotations

import logging
import os
from pathlib import Path

import torch
import torch.nn as nn

from .config import TransformerConfig, ViTConfig
from .layers import Transformer

logger = logging.getLogger("vitef")

VIT_SIZES = {
    "base": dict(emb_dim=768, n_heads=12, n_layers=12, ffn_dim=3072),
    "large": dict(emb_dim=1024, n_heads=16, n_layers=24, ffn_dim=4096),
    "huge": dict(emb_dim=1280, n_heads=16, n_layers=32, ffn_dim=5120),
}


class ViT(nn.Module):
    def __init__(self, vit_config: ViTConfig):
        super().__init__()
        self.model_name = f"vit-{vit_config.model_name.lower()}-patch{vit_config.patch_size}-{vit_config.image_dim[-1]}"
        if vit_config.in21k:
            self.model_name += "-in21k"
        args = dict(VIT_SIZES[vit_config.model_name])
        args.update(
            image_dim=vit_config.image_dim, patch_type="computer_vision", image_patch="hybrid", patch_size=vit_config.patch_size,
            emb_type="linear", pos_emb=True, freeze_pos=False, emb_dropout=0.0, attn_bias=True, attn_dropout=0.0, flash=False,
            causal=False, activation="gelu", ffn_bias=True, ffn_dropout=0.0, norm="layer", norm_bias=True, norm_eps=1e-12,
            pre_norm=True, cls_token=True, output_type="classification", weight_tying=False, output_dropout=0.0,
            n_classes=1000 if not vit_config.in21k else 2,
        )
        config = TransformerConfig(**args)
        self.model = Transformer(config)
        self.config = config
        # instance attributes, as in the reference: ViT.__call__ hooks fire, the inner module's do not
        self.forward = self.model.forward
        self.get_decomposition = self.model.get_decomposition
        self.get_probes = self.model.get_probes
        self.get_pooled_probes = self.model.get_pooled_probes  # on-device pooling for linear probing (not in the reference)

        if vit_config.pretrained:
            self.save_dir = vit_config.save_dir
            self.load_pretrained_weights()
        if vit_config.finetuning:
            self.config.n_classes = vit_config.n_classes
            self.set_finetuning_mode()
            logger.info(f"Initialize new classification head with {self.config.n_classes} classes for finetuning.")

    def load_pretrained_weights(self) -> None:
        """Load ``<save_dir>/<model_name>.pt`` (vit.py:214-225). The reference downloads from HuggingFace when the file is
        missing (239-303); this offline build cannot, and silently finetuning / analysing a RANDOM model in its place
        would produce plausible but meaningless numbers, so a missing file raises. Random-init runs (the BASELINE
        configs) pass ``pretrained=False``."""
        path = Path(self.save_dir) / f"{self.model_name}.pt" if self.save_dir is not None else None
        if path is None or not os.path.exists(path):
            raise FileNotFoundError(f"pretrained=True but there are no local weights for {self.model_name} at {path} (this build does not "
                                    f"download from HuggingFace); place the converted state_dict there or pass pretrained=False")
        logger.info(f"Loading {self.model_name} model from {path}")
        self.model.load_state_dict(torch.load(path))

    def set_finetuning_mode(self) -> None:
        head = self.model.output.output_layer
        old = head.output.weight
        head.output = nn.Linear(self.config.emb_dim, self.config.n_classes).to(device=old.device)
