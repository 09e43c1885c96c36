"""`with Quantize(model, optim, cfg.quant) as q: ...; q.convert()` (reference: pipeline/quant/context.py).

KMeans : weights are re-clustered at weight load before every fused forward (kmeans.py).
QAT    : torch's default 'fbgemm' QAT qconfig, as torch.quantization.prepare_qat applies it to the reference model.
         Weights: each nn.Linear weight is fake-quantised at weight load (per-output-channel symmetric int8,
         running min/max with averaging constant 0.01); the straight-through gradient goes to the fp32 master.
         Activations: the OUTPUT of every nn.Linear goes through FusedMovingAvgObsFakeQuantize (per-tensor affine
         quint8 with reduce_range = [0, 127], moving-average min/max) — sirenb200_set_act_quant, on the fp32
         engine (`activations=False` in the quant config keeps the weights-only tensor-core variant).
         `convert()` freezes the observers and returns the model in eval mode with, on every Linear,
         `weight_codes` (int8), `weight_scales`, `act_scale`, `act_zero_point` — the tensors of torch's converted
         int8 module; its forward computes what the int8 graph computes (fake-quantised values are exactly
         representable, so quantise -> int8 GEMM -> requantise equals the float evaluation up to fp32 summation).
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

    # ------------------------------------------------------------------ QAT
    def _prepare_QAT(self):
        if not hasattr(self.model, "_weight_transforms"):
            raise _lib.SirenB200Error("QAT needs the fused Siren model")
        qc = dict(self.quant_conf)
        self._observers = {}
        self._targets = [(n, m) for n, m in self.model.named_modules() if isinstance(m, nn.Linear)]
        for name, _ in self._targets:
            self._observers[name] = _QATWeightObserver()
        self.model._weight_transforms.append(self._fake_quant_weights)
        self._activations = bool(qc.get("activations", True))
        if self._activations:
            dev = next(self.model.parameters()).device
            state = torch.tensor([[float("inf"), float("-inf"), 1.0, 0.0]] * self.model.depth, dtype=torch.float32,
                                 device=dev)
            self.model.precision = "fp32"  # activation fake-quant lives on the fp32 engine
            self.model._act_quant = {"state": state, "averaging_constant": 0.01, "qmin": 0, "qmax": 127}

    def _fake_quant_weights(self, model):
        for name, m in self._targets:
            w = m.weight.data
            if not getattr(self, "_frozen", False):
                # torch's FusedMovingAvgObsFakeQuantize updates its observer on EVERY forward, eval_epoch's included
                # (observer_enabled does not follow module.training); convert() is what freezes it
                lo, hi = self._observers[name].update(w)
            else:
                ob = self._observers[name]
                lo, hi = (ob.min_val, ob.max_val) if ob.min_val is not None else torch.aminmax(w, dim=1)
            codes, scales, wq = _engine.fakequant_per_channel(w, lo, hi)
            model._param_override[m.weight] = wq  # kernels read wq; gradients go to the master weight
            m.weight_codes, m.weight_scales = codes, scales

    def _convert_QAT(self):
        """context.py:28-29 (torch.quantization.convert(model.eval())): freeze everything and expose the int8
        module's tensors."""
        self.model.eval()
        self._frozen = True
        self._fake_quant_weights(self.model)
        self.model._weight_transforms.remove(self._fake_quant_weights)
        for _, m in self._targets:
            m.weight.data = self.model._param_override.pop(m.weight)
        if self._activations:
            aq = self.model._act_quant
            aq["frozen"] = True
            for i, layer in enumerate(self.model.layers):
                layer.linear.act_scale = aq["state"][i, 2].clone()
                layer.linear.act_zero_point = aq["state"][i, 3].to(torch.int32)
        return self.model
