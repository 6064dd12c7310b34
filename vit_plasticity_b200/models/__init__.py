from .config import TransformerConfig, ViTConfig
from .factory import build_model
from .layers import Transformer
from .vit import ViT

__all__ = ["Transformer", "TransformerConfig", "ViT", "ViTConfig", "build_model"]
