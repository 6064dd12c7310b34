"""Configuration dataclasses, field-compatible with the reference so YAML / dict configs carry over unchanged.

TransformerConfig mirrors src/vitef/models/transformer/architecture.py:48-123 and ViTConfig mirrors
src/vitef/models/vit.py:39-80 (same field names and defaults, same kwargs-filtering constructor, same
post-init defaulting). ``build_with_type_check`` follows src/vitef/utils.py:44-99 for the cases the ViT path uses.
"""

from __future__ import annotations

import logging
from dataclasses import dataclass, fields, is_dataclass
from pathlib import Path
from typing import Any, Union, get_args, get_origin

logger = logging.getLogger("vitef")

MODEL_DIR = Path(__file__).resolve().parents[2] / "checkpoints"


@dataclass
class TransformerConfig:
    # data
    image_dim: tuple = (3, 224, 224)
    length: int = 512
    # patching
    patch_type: str | None = None
    image_patch: str = "hybrid"
    patch_size: int = 16
    stride: int = 8
    # embedding
    vocab_size: int = -1
    emb_type: str = "dict"
    emb_dim: int = -1
    pos_emb: bool = True
    freeze_pos: bool = False
    seq_len: int = -1
    emb_dropout: float | None = None
    # attention
    n_heads: int = -1
    attn_bias: bool = False
    attn_dropout: float | None = None
    flash: bool = False
    causal: bool = False
    # feed-forward
    activation: str = "gelu"
    ffn_dim: int | None = None
    ffn_bias: bool = False
    ffn_dropout: float | None = None
    # block
    norm: str = "layer"
    norm_bias: bool = False
    norm_eps: float = 1e-5
    pre_norm: bool = True
    # stack
    n_layers: int = -1
    dropout: float = 0.0
    # task head
    cls_token: bool = False
    output_type: str = "sequence_to_sequence"
    weight_tying: bool = True
    output_dropout: float | None = None
    n_classes: int = -1
    forecasting_horizon: int = -1

    def __init__(self, **kwargs):
        known = {f.name: f.default for f in fields(self)}
        for name, default in known.items():
            setattr(self, name, kwargs.get(name, default))
        self.__post_init__()

    def __post_init__(self):
        if self.ffn_dim is None:
            self.ffn_dim = 4 * self.emb_dim
        if self.flash is None:
            self.flash = True
        for name in ("emb_dropout", "attn_dropout", "ffn_dropout", "output_dropout"):
            if getattr(self, name) is None:
                setattr(self, name, self.dropout)
        if isinstance(self.image_dim, list):
            self.image_dim = tuple(self.image_dim)


@dataclass
class ViTConfig:
    model_name: str = "base"
    pretrained: bool = False
    in21k: bool = False
    save_dir: str = None
    patch_size: int = 16
    image_dim: tuple = (3, 224, 224)
    finetuning: bool = False
    n_classes: int = 1000

    def __init__(self, **kwargs):
        known = {f.name: f.default for f in fields(self)}
        for name, default in known.items():
            setattr(self, name, kwargs.get(name, default))
        self.__post_init__()

    def __post_init__(self):
        if self.save_dir is None:
            self.save_dir = MODEL_DIR / "vit"
        if isinstance(self.image_dim, list):
            self.image_dim = tuple(self.image_dim)


def build_with_type_check(object_type: Any, data: Any) -> Any:
    """Build a (possibly nested) dataclass from a dict, consuming known keys and warning about the rest."""
    if data is None or object_type is Any:
        return data
    if is_dataclass(object_type):
        values = {}
        for f in fields(object_type):
            if f.init and f.name in data:
                values[f.name] = build_with_type_check(f.type if not isinstance(f.type, str) else Any, data.pop(f.name))
        for leftover in data:
            logger.warning(f"Field '{leftover}' ignored when initializing {object_type}.")
        return object_type(**values)
    origin, args = get_origin(object_type), get_args(object_type)
    if origin is list and len(args) == 1:
        return [build_with_type_check(args[0], v) for v in data]
    if origin is dict and len(args) == 2:
        return {build_with_type_check(args[0], k): build_with_type_check(args[1], v) for k, v in data.items()}
    if origin is Union:
        for a in args:
            try:
                return build_with_type_check(a, data)
            except (TypeError, ValueError):
                continue
        return data
    try:
        return object_type(data)
    except (TypeError, ValueError):
        return data
