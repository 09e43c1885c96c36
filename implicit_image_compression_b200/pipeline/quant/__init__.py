from .context import Quantize
from .kmeans import KmeansQuant

__all__ = ["Quantize", "KmeansQuant"]
