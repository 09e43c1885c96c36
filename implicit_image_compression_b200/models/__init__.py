"""Model registry (reference: implicit_image/models/__init__.py:5).  "siren" is the tensor-core hot path; "fourier"
 runs on the library's fp32 path; "wavelet_siren" is out of scope."""
from .fourier import FourierNet
from .siren import Siren, SineLayer

registry = {"siren": Siren, "fourier": FourierNet}

__all__ = ["registry", "Siren", "SineLayer", "FourierNet"]
