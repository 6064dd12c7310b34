"""``build_model`` with the reference's contract (src/vitef/models/utils.py:25-85): pops ``implementation`` from the
caller's dict, builds the typed config (unknown keys are warned about and ignored), instantiates, moves to
``device`` and optionally returns ``asdict(config)``. Only the ViT path ("vit", "transformer") is implemented."""

from __future__ import annotations

from dataclasses import asdict
from typing import Any

import torch
import torch.nn as nn

from .config import TransformerConfig, ViTConfig, build_with_type_check

DEVICE = "cuda" if torch.cuda.is_available() else "cpu"


def build_model(config: dict[str, Any], device: str = DEVICE, return_config: bool = False) -> nn.Module:
    implementation = config.pop("implementation", "vit")
    match implementation.lower():
        case "transformer":
            from .layers import Transformer

            model_type, config_obj = Transformer, build_with_type_check(TransformerConfig, config)
        case "vit":
            from .vit import ViT

            model_type, config_obj = ViT, build_with_type_check(ViTConfig, config)
        case "gpt2" | "patchtst":
            raise NotImplementedError(f"'{implementation}' is outside the ViT hot path this package implements (SURVEY.md section 2, row 15)")
        case _:
            raise ValueError(f"Implementation {implementation} not found.")
    model = model_type(config_obj).to(device=device)
    if return_config:
        return model, asdict(config_obj)
    return model
