"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference (varun19299/implicit-image-compression)
from /root/reference with stub modules for third-party imports that are not installed here
(SURVEY.md §8c recipe).  Used by tools/make_golden.py (fixture generation, this container only) and by
tests that pin oracle/siren_oracle.py against the reference when /root/reference exists.

Nothing under the product package may import this module.  /root/reference does not exist on the GPU
box: `available()` is False there and callers must skip.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SIRENB200_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "implicit_image"))


class AttrDict(dict):
    """Minimal stand-in for omegaconf.DictConfig: attribute access + .get()."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v

    def __setattr__(self, k, v):
        self[k] = v


def _scatter_mean(src, index, dim=0):
    """torch_scatter.scatter_mean restated (reference call site: quant/kmeans_helper.py:89):
    per-cluster mean, output length index.max()+1, empty clusters -> 0."""
    import torch

    assert dim == 0
    n = int(index.max().item()) + 1
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    cnt = torch.zeros(n, dtype=src.dtype, device=src.device)
    out.index_add_(0, index, src)
    cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    return out / cnt.clamp(min=1).reshape((n,) + (1,) * (src.dim() - 1))


def _install_stubs():
    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _TensorType:
        def __class_getitem__(cls, item):
            return cls

    for name in ("omegaconf", "torch_optimizer", "kornia", "pytorch_wavelets", "matplotlib",
                 "matplotlib.pyplot", "torchtyping", "torch_scatter", "zstandard", "hydra", "colorlog"):
        try:
            importlib.import_module(name)
        except Exception:
            pass
    if "omegaconf" not in sys.modules:
        mod("omegaconf", DictConfig=dict, OmegaConf=type("OmegaConf", (), {}))
    if "torch_optimizer" not in sys.modules:
        mod("torch_optimizer", Shampoo=object)
    if "kornia" not in sys.modules:
        mod("kornia")
    if "pytorch_wavelets" not in sys.modules:
        mod("pytorch_wavelets", DWTInverse=object, DWTForward=object)
    if "matplotlib" not in sys.modules:
        mpl = mod("matplotlib")
        mpl.pyplot = mod("matplotlib.pyplot")
    if "torchtyping" not in sys.modules:
        mod("torchtyping", TensorType=_TensorType)
    if "torch_scatter" not in sys.modules:
        mod("torch_scatter", scatter_mean=_scatter_mean)
    if "zstandard" not in sys.modules:  # only the 'zstd' stream needs it; the 'plain' stream is pure Python
        mod("zstandard")


_cache = {}


def load():
    """Returns a namespace with the reference's hot-path symbols."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    from implicit_image.models import registry as model_registry  # noqa
    from implicit_image.data import get_grid  # noqa
    from implicit_image.utils import train_helper  # noqa
    from implicit_image.pipeline.masking import Masking  # noqa
    from implicit_image.pipeline.masking.funcs.decay import registry as decay_registry  # noqa
    from implicit_image.pipeline.quant.context import Quantize  # noqa
    from implicit_image.pipeline.quant.kmeans import KmeansQuant  # noqa

    ns.model_registry = model_registry
    ns.get_grid = get_grid
    ns.train_epoch = train_helper.train_epoch
    ns.eval_epoch = train_helper.eval_epoch
    ns.get_optimizer_lr_scheduler = train_helper.get_optimizer_lr_scheduler
    ns.setup_mask = train_helper.setup_mask
    ns.Masking = Masking
    ns.decay_registry = decay_registry
    ns.Quantize = Quantize
    ns.KmeansQuant = KmeansQuant
    ns.AttrDict = AttrDict
    _cache["ns"] = ns
    return ns


def importlib_import(name):
    """Import another module of the (already loaded, stubbed) reference package, e.g. the entropy coder."""
    import importlib
    load()
    return importlib.import_module(name)
