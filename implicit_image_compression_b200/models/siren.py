"""SIREN coordinate MLP whose forward / backward run in libsirenb200 (sm_100a CUDA).

Drop-in for the reference model (implicit_image/models/siren.py:10-134): same constructor keywords, same
parameter names (`layers.{i}.linear.weight|bias`, fp32, [out, in]), `nn.Linear` sub-modules (the masking and
k-means code filter on them), same RNG consumption at init so `torch.manual_seed(s); Siren(...)` draws the
same weights.  The per-layer math is NOT executed by these modules: `Siren.forward` hands the parameter
pointers to the engine, which runs the fused kernels (in-kernel coordinates, tcgen05 GEMM + sin epilogue).
"""
import numpy as np
import torch
from torch import nn

from .. import _lib
from ..engine import SirenEngine, default_precision


class SineLayer(nn.Module):
    """Parameter holder for one layer: z = x W^T + b, a = sin(omega_0 z) (siren.py:56-68)."""

    def __init__(self, in_features, out_features, has_bias=True, is_first=False, omega_0=30.0,
                 no_activation=False, simulate_quantization=False):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.has_bias, self.is_first = has_bias, is_first
        self.omega_0, self.no_activation = omega_0, no_activation
        self.simulate_quantization = simulate_quantization
        self.linear = nn.Linear(in_features, out_features, bias=has_bias)
        self.quant = torch.quantization.QuantStub()
        self.dequant = torch.quantization.DeQuantStub()
        self.init_weights()

    @torch.no_grad()
    def init_weights(self):
        # siren.py:45-54: U(-1/in, 1/in) for the first layer, U(-sqrt(6/in)/omega, +) otherwise; the bias
        # keeps nn.Linear's default init.
        bound = 1 / self.in_features if self.is_first else np.sqrt(6 / self.in_features) / self.omega_0
        self.linear.weight.uniform_(-bound, bound)
        self.linear.scaler = bound

    def forward(self, x):
        raise _lib.SirenB200Error("SineLayer is a parameter holder; call Siren.forward (fused CUDA path)")


class _SirenFunction(torch.autograd.Function):
    """autograd bridge: forward -> sirenb200_forward, backward -> sirenb200_backward."""

    @staticmethod
    def forward(ctx, model, grid, *params):
        engine = model.engine_for(grid)
        kparams = model.kernel_parameters()
        pred = engine.forward(kparams)
        ctx.engine, ctx.generation, ctx.kparams = engine, engine.generation, kparams
        ctx.save_for_backward(*params)
        return pred

    @staticmethod
    def backward(ctx, dpred):
        engine = ctx.engine
        if engine.generation != ctx.generation:
            raise _lib.SirenB200Error("stale activations: another forward ran on this model before backward")
        grads = [torch.empty_like(p) for p in ctx.kparams]
        engine.backward(ctx.kparams, dpred.contiguous(), grads)
        return (None, None, *grads)


class Siren(nn.Module):
    def __init__(self, input_size=2, output_size=3, depth=8, hidden_size=128, first_omega_0=50.0,
                 hidden_omega_0=50.0, outermost_linear=True, simulate_quantization=False,
                 small_dense_density=1.0, width=None, precision=None, **kwargs):
        super().__init__()
        if width is not None:  # README-era alias of mlp.hidden_size (SURVEY.md §0 naming drift)
            hidden_size = width
        hidden_size = int(hidden_size * np.sqrt(small_dense_density))  # siren.py:88
        if input_size != 2:
            raise _lib.SirenB200Error("the B200 path supports 2-D pixel coordinates only (input_size=2)")
        layers = [SineLayer(input_size, hidden_size, is_first=True, omega_0=first_omega_0,
                            simulate_quantization=simulate_quantization)]
        for _ in range(depth - 2):
            layers.append(SineLayer(hidden_size, hidden_size, omega_0=hidden_omega_0,
                                    simulate_quantization=simulate_quantization))
        layers.append(SineLayer(hidden_size, output_size, omega_0=hidden_omega_0,
                                no_activation=outermost_linear,
                                simulate_quantization=simulate_quantization))
        self.simulate_quantization = simulate_quantization
        self.layers = nn.Sequential(*layers)
        self.depth, self.hidden_size, self.output_size = depth, hidden_size, output_size
        self.first_omega_0, self.hidden_omega_0 = float(first_omega_0), float(hidden_omega_0)
        self.outermost_linear = bool(outermost_linear)
        self.precision = precision  # None = automatic ("f16tc" when hidden in {128, 256}, else "fp32")
        self._engines = {}
        self._weight_transforms = []  # callables run before every forward ("weight load" hooks)
        self._param_override = {}     # param -> tensor the kernels read instead (fake-quantised weights)
        self._post_backward = []      # callables run after the fused backward filled param.grad

    # --------------------------------------------------------------------------------------------
    def hot_parameters(self):
        """[w0, b0, w1, b1, ...] in model.parameters() order."""
        out = []
        for layer in self.layers:
            out += [layer.linear.weight, layer.linear.bias]
        return out

    def kernel_parameters(self):
        """Tensors whose pointers go to the kernels: the parameters, or their weight-load overrides."""
        return [self._param_override.get(p, p.data) for p in self.hot_parameters()]

    def _precision_code(self):
        if self.precision in (None, "auto"):
            return default_precision(self.hidden_size, self.depth)
        return {"fp32": _lib.PREC_FP32, "f16tc": _lib.PREC_F16TC}[self.precision]

    def engine_for(self, grid, row_begin=0, row_end=None, height=None):
        """Engine (workspace) for this grid's geometry; cached per (H, W, rows, precision, device)."""
        _lib.require_cuda(grid, "grid")
        h, w = int(grid.shape[0]), int(grid.shape[1])
        height = h if height is None else height
        row_end = (row_begin + h) if row_end is None else row_end
        key = (height, w, row_begin, row_end, self._precision_code(), grid.device.index)
        eng = self._engines.get(key)
        if eng is None:
            eng = SirenEngine(self.depth, self.hidden_size, self.first_omega_0, self.hidden_omega_0,
                              self.outermost_linear, self.output_size, height, w, row_begin, row_end,
                              self._precision_code(), grid.device)
            self._engines[key] = eng
        eng.bind_grid(grid)
        aq = self.__dict__.get("_act_quant")
        if aq is not None:  # QAT activation observers (pipeline/quant/context.py).  torch's observers keep updating
            # in eval mode too (observer_enabled does not follow module.training); only convert() freezes them
            eng.set_act_quant(aq["state"], not aq.get("frozen", False), aq["averaging_constant"],
                              aq["qmin"], aq["qmax"])
        elif getattr(eng, "_act_state", None) is not None:
            eng.set_act_quant(None, False)
        return eng

    def run_weight_transforms(self):
        """Weight-load hooks: k-means re-clustering / fake quantisation overwrite weight.data before the
        kernels stage the weights (reference: forward_pre_hook, quant/kmeans.py:48,65-71)."""
        for fn in self._weight_transforms:
            fn(self)

    def forward(self, grid):
        """grid [h, w, 2] in [0,1] -> [h, w, output_size] (siren.py:123-134)."""
        _lib.require_cuda(grid, "grid")
        self.run_weight_transforms()
        params = self.hot_parameters()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _SirenFunction.apply(self, grid, *params)
        return self.engine_for(grid).forward(self.kernel_parameters())

    def __deepcopy__(self, memo):
        # engines own device workspaces and ctypes handles: never copy them (compress.py:174 deepcopy)
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_engines":
                new.__dict__[k] = {}
            elif k == "_weight_transforms":
                new.__dict__[k] = []
            elif k == "_param_override":
                new.__dict__[k] = {}
            elif k == "_post_backward":
                new.__dict__[k] = []
            elif k == "_step_fitters":
                new.__dict__[k] = {}
            elif k == "_act_quant":
                continue
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new
