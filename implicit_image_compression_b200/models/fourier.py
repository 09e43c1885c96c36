"""FourierNet whose forward / backward run in libsirenb200 (reference: implicit_image/models/fourier.py:8-69).

Same constructor keywords, parameter names (`encoding.B`, `layers.{0,2,4,...}.weight|bias` — the nn.Sequential of the
reference interleaves nn.ReLU modules, so the Linear layers sit at even indices) and RNG consumption (the Linear
layers are created first, then `torch.randn` for the encoding) as the reference.  The math — Fourier-feature
encoding of the RAW [0, 1] coordinates, Linear + ReLU layers, Linear + Sigmoid output — runs on the library's fp32
CUDA-core path (model family 1 of sirenb200_create); there is no tcgen05 variant of it yet."""
import numpy as np
import torch
from torch import nn

from .. import _lib
from ..engine import SirenEngine
from .siren import Siren, _SirenFunction


class Encoding(nn.Module):
    """Parameter holder of the random Fourier features (fourier.py:8-25)."""

    def __init__(self, input_size=2, map_size=256, map_scale=10.0):
        super().__init__()
        assert map_size % 2 == 0, "Need even map size"
        self.B = nn.Parameter(torch.randn(input_size, map_size // 2) * map_scale, requires_grad=False)

    def forward(self, x):
        raise _lib.SirenB200Error("Encoding is a parameter holder; call FourierNet.forward (fused CUDA path)")


class FourierNet(Siren):
    def __init__(self, input_size=2, output_size=3, depth=8, hidden_size=128, map_size=128, map_scale=10.0,
                 small_dense_density=1.0, precision=None, **kwargs):
        nn.Module.__init__(self)
        if input_size != 2:
            raise _lib.SirenB200Error("the B200 path supports 2-D pixel coordinates only (input_size=2)")
        if precision not in (None, "auto", "fp32"):
            raise _lib.SirenB200Error("FourierNet runs on the fp32 path")
        hidden_size = int(hidden_size * np.sqrt(small_dense_density))  # fourier.py:42
        layers = [nn.Linear(map_size, hidden_size), nn.ReLU(inplace=True)]
        for _ in range(depth - 3):
            layers += [nn.Linear(hidden_size, hidden_size), nn.ReLU(inplace=True)]
        layers += [nn.Linear(hidden_size, output_size), nn.Sigmoid()]
        self.encoding = Encoding(input_size, map_size, map_scale)
        self.layers = nn.Sequential(*layers)
        self.depth, self.hidden_size, self.output_size, self.map_size = depth, hidden_size, output_size, map_size
        self.first_omega_0 = self.hidden_omega_0 = 1.0
        self.outermost_linear = True
        self.simulate_quantization = False
        self.precision = "fp32"
        self._engines = {}
        self._weight_transforms = []
        self._param_override = {}
        self._post_backward = []

    def hot_parameters(self):
        out = []
        for m in self.layers:
            if isinstance(m, nn.Linear):
                out += [m.weight, m.bias]
        return out

    def _precision_code(self):
        return _lib.PREC_FP32

    def engine_for(self, grid, row_begin=0, row_end=None, height=None):
        _lib.require_cuda(grid, "grid")
        h, w = int(grid.shape[0]), int(grid.shape[1])
        height = h if height is None else height
        row_end = (row_begin + h) if row_end is None else row_end
        key = (height, w, row_begin, row_end, grid.device.index)
        eng = self._engines.get(key)
        if eng is None:
            eng = SirenEngine(self.depth, self.hidden_size, 1.0, 1.0, True, self.output_size, height, w, row_begin,
                              row_end, _lib.PREC_FP32, grid.device, model_kind=1, map_size=self.map_size)
            self._engines[key] = eng
        eng.bind_grid(grid)
        eng.set_fourier_encoding(self.encoding.B.data)
        return eng

    def forward(self, grid):
        """grid [h, w, 2] in [0,1] -> [h, w, output_size] (fourier.py:58-69)."""
        _lib.require_cuda(grid, "grid")
        self.run_weight_transforms()
        params = self.hot_parameters()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _SirenFunction.apply(self, grid, *params)
        return self.engine_for(grid).forward(self.kernel_parameters())
