"""Model registry (reference: implicit_image/models/__init__.py:5).  Only "siren" is on the B200 hot path;
"fourier" / "wavelet_siren" are catalogued as out of scope."""
from .siren import Siren, SineLayer

registry = {"siren": Siren}

__all__ = ["registry", "Siren", "SineLayer"]
