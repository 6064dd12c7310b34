"""B200-native (sm_100a) implementation of the ViT hot path of ambroiseodt/vit-plasticity.

Public surface (mirrors ``vitef.models``): ``build_model``, ``ViT``, ``ViTConfig``, ``Transformer``,
``TransformerConfig``; plus ``plasticity`` (fused estimator), ``finetune`` (train step, freeze_model) and
``distributed`` (data-parallel gradient all-reduce). All compute goes through ``libvitb200.so`` (``_lib``).
"""

from .models import Transformer, TransformerConfig, ViT, ViTConfig, build_model

__all__ = ["Transformer", "TransformerConfig", "ViT", "ViTConfig", "build_model"]
__version__ = "0.1.0"
