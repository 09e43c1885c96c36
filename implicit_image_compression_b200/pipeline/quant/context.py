"""`with Quantize(model, optim, cfg.quant) as q: ...; q.convert()` (reference: pipeline/quant/context.py).

KMeans : weights are re-clustered at weight load before every fused forward (kmeans.py).
QAT    : weights-only quantisation-aware training.  Each nn.Linear weight is fake-quantised at weight load
         (per-output-channel symmetric int8, running min/max with averaging constant 0.01, exactly the
         weight half of torch's default 'fbgemm' QAT qconfig) and the straight-through gradient is applied
         to the fp32 master weight.  Activation observers and the int8 `torch.quantization.convert`
         inference graph are "next" (SURVEY.md §8f rank 2); `convert()` returns the model with dequantised
         weights plus `weight_codes` (int8) and `weight_scales` on every Linear.
"""
import torch
from torch import nn

from ... import _lib
from ... import engine as _engine
from .kmeans import KmeansQuant


class _QATWeightObserver:
    """MovingAveragePerChannelMinMaxObserver (ch_axis 0, averaging_constant 0.01) for one weight."""

    def __init__(self, averaging_constant=0.01):
        self.c = averaging_constant
        self.min_val = None
        self.max_val = None

    def update(self, w):
        lo, hi = torch.aminmax(w, dim=1)
        if self.min_val is None:
            self.min_val, self.max_val = lo, hi
        else:
            self.min_val = self.min_val + self.c * (lo - self.min_val)
            self.max_val = self.max_val + self.c * (hi - self.max_val)
        return self.min_val, self.max_val


class Quantize:
    def __init__(self, model, optim, quant_conf):
        self.model, self.optim, self.quant_conf = model, optim, quant_conf

    def _name(self):
        qc = self.quant_conf
        return qc["name"] if isinstance(qc, dict) else qc.name

    def __enter__(self):
        getattr(self, f"_prepare_{self._name()}")()
        return self

    def __exit__(self, exc_type, exc, tb):
        return False

    def convert(self):
        return getattr(self, f"_convert_{self._name()}")()

    # ------------------------------------------------------------------ KMeans
    def _prepare_KMeans(self):
        qc = dict(self.quant_conf)
        skip_ll = qc.get("skip_ll", ["layers.0.linear", "layers.7.linear"])
        self.compress = KmeansQuant(self.model, self.optim, bits=qc["bits"], skip_ll=skip_ll)

    def _convert_KMeans(self):
        self.compress.update_weights()
        return self.model

    # ------------------------------------------------------------------ QAT (weights only)
    def _prepare_QAT(self):
        if not hasattr(self.model, "_weight_transforms"):
            raise _lib.SirenB200Error("QAT needs the fused Siren model")
        self._observers = {}
        self._targets = [(n, m) for n, m in self.model.named_modules() if isinstance(m, nn.Linear)]
        for name, _ in self._targets:
            self._observers[name] = _QATWeightObserver()
        self.model._weight_transforms.append(self._fake_quant_weights)

    def _fake_quant_weights(self, model):
        for name, m in self._targets:
            w = m.weight.data
            if model.training:
                lo, hi = self._observers[name].update(w)
            else:
                ob = self._observers[name]
                lo, hi = (ob.min_val, ob.max_val) if ob.min_val is not None else torch.aminmax(w, dim=1)
            codes, scales, wq = _engine.fakequant_per_channel(w, lo, hi)
            model._param_override[m.weight] = wq  # kernels read wq; gradients go to the master weight
            m.weight_codes, m.weight_scales = codes, scales

    def _convert_QAT(self):
        self.model.eval()
        self._fake_quant_weights(self.model)
        self.model._weight_transforms.remove(self._fake_quant_weights)
        for _, m in self._targets:
            m.weight.data = self.model._param_override.pop(m.weight)
        return self.model
