from .core import Masking, LayerStats

__all__ = ["Masking", "LayerStats"]
