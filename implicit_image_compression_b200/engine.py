"""SirenEngine: one libsirenb200 handle (workspace) bound to a model shape, an image geometry and a row
range.  PyTorch owns parameters / gradients / optimizer state; the engine receives raw pointers per call
(the reference rebinds weight.data) and launches on torch's current stream.
"""
import ctypes

import torch

from . import _lib


def default_precision(hidden, depth=3):
    """fp16 tensor-core path for hidden widths 64..512 (any width: the library zero-pads it to the 128 / 256 / 512
    columns its tcgen05 kernels are built for — models/siren.py:88 produces widths like 114), else the fp32
    CUDA-core path (tiny plumbing models, hidden > 512, depth 2)."""
    if hidden in (128, 256, 512) or (64 < hidden <= 512 and depth >= 3):
        return _lib.PREC_F16TC
    return _lib.PREC_FP32


class SirenEngine:
    def __init__(self, depth, hidden, first_omega, hidden_omega, outermost_linear, out_features,
                 height, width, row_begin=0, row_end=None, precision=None, device=None, model_kind=0, map_size=0):
        if not torch.cuda.is_available():
            raise _lib.SirenB200Error("no CUDA device: siren-b200 needs an sm_100a GPU (no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        if self.device.type != "cuda":
            raise _lib.SirenB200Error(f"engine device must be CUDA, got {self.device}")
        row_end = height if row_end is None else row_end
        if precision is None:
            precision = default_precision(hidden, depth)
        self.cfg = _lib.Config(depth=depth, hidden=hidden, in_features=2, out_features=out_features,
                               first_omega=float(first_omega), hidden_omega=float(hidden_omega),
                               outermost_linear=int(bool(outermost_linear)), height=height, width=width,
                               row_begin=row_begin, row_end=row_end, precision=precision)
        self.cfg.reserved[0], self.cfg.reserved[1] = int(model_kind), int(map_size)
        self.model_kind = int(model_kind)
        self.num_linear = depth - 1 if model_kind == 1 else depth  # FourierNet: depth - 1 nn.Linear layers
        self.depth, self.hidden, self.out_features = depth, hidden, out_features
        self.height, self.width = height, width
        self.row_begin, self.row_end = row_begin, row_end
        self.rows = row_end - row_begin
        self.precision = precision
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sirenb200_create(ctypes.byref(self.cfg), ctypes.byref(self.handle)))
        self._grid_keepalive = None
        self._grid_key = None
        self.generation = 0

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                self.lib.sirenb200_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ grid
    def bind_grid(self, grid):
        """grid: [rows, W, 2] (this engine's rows) or the full [H, W, 2] grid, values in [0, 1].
        If the grid is an outer product of its first column / first row (data.get_grid), only those two
        tables are kept and coordinates are generated in-kernel; otherwise the explicit tensor is used."""
        _lib.require_cuda(grid, "grid")
        key = (grid.data_ptr(), tuple(grid.shape), grid._version)
        if key == self._grid_key:
            return
        g = grid
        if g.shape[0] == self.height and self.rows != self.height:
            g = g[self.row_begin:self.row_end]
        if tuple(g.shape) != (self.rows, self.width, 2):
            raise _lib.SirenB200Error(f"grid shape {tuple(grid.shape)} does not match engine "
                                      f"({self.rows}x{self.width}x2)")
        g = g.to(torch.float32)
        lin_h = g[:, 0, 0].contiguous()
        lin_w = g[0, :, 1].contiguous()
        separable = bool(((g[..., 0] == lin_h[:, None]) & (g[..., 1] == lin_w[None, :])).all().item())
        if separable:
            # the LUT is indexed by absolute image row
            full_h = torch.zeros(self.height, dtype=torch.float32, device=g.device)
            full_h[self.row_begin:self.row_end] = lin_h
            self._grid_keepalive = (full_h, lin_w)
            _lib.check(self.lib.sirenb200_set_grid_lut(self.handle, full_h.data_ptr(), lin_w.data_ptr()))
        else:
            gc = g.contiguous()
            self._grid_keepalive = (gc,)
            _lib.check(self.lib.sirenb200_set_grid_coords(self.handle, gc.data_ptr()))
        self._grid_key = key

    # ------------------------------------------------------------------ compute
    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def _check_params(self, params):
        if len(params) != 2 * self.num_linear:
            raise _lib.SirenB200Error(f"expected {2 * self.num_linear} parameter tensors, got {len(params)}")
        for p in params:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise _lib.SirenB200Error("parameters must be contiguous fp32 CUDA tensors "
                                          f"(got {p.dtype} on {p.device})")

    def forward(self, params, out=None):
        """Siren.forward for this engine's rows -> [rows, W, C] fp32."""
        self._check_params(params)
        if out is None:
            out = torch.empty(self.rows, self.width, self.out_features, dtype=torch.float32,
                              device=self.device)
        _lib.check(self.lib.sirenb200_forward(self.handle, _lib.ptr_array(params), out.data_ptr(),
                                              self._stream()))
        self.generation += 1
        return out

    def forward_backward(self, params, img, grads, stats=None, loss_scale=1.0):
        """forward + MSE + backward.  grads[i] are overwritten.  Returns the device stats tensor
        [sum_sq_err, loss, nonfinite_flag, 0]."""
        self._check_params(params)
        _lib.require_cuda(img, "img")
        if stats is None:
            stats = torch.empty(4, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.sirenb200_forward_backward(
            self.handle, _lib.ptr_array(params), img.data_ptr(), float(loss_scale),
            _lib.ptr_array(grads), stats.data_ptr(), self._stream()))
        self.generation += 1
        return stats

    def backward(self, params, dpred, grads):
        self._check_params(params)
        _lib.check(self.lib.sirenb200_backward(self.handle, _lib.ptr_array(params), dpred.data_ptr(),
                                               _lib.ptr_array(grads), self._stream()))

    def set_act_quant(self, state, training, averaging_const=0.01, qmin=0, qmax=127):
        """QAT activation fake-quant on every layer's pre-activation (fp32 engines).  state: [depth, 4] fp32 CUDA
        tensor {running min, running max, scale, zero point}; None switches it off."""
        if state is None:
            _lib.check(self.lib.sirenb200_set_act_quant(self.handle, None, 0, 0, 0.0, 0, 1))
            self._act_state = None
            return
        _lib.require_cuda(state, "state")
        if state.dtype != torch.float32 or tuple(state.shape) != (self.num_linear, 4) or not state.is_contiguous():
            raise _lib.SirenB200Error("activation-quant state must be a contiguous [depth, 4] fp32 tensor")
        self._act_state = state  # keep-alive
        _lib.check(self.lib.sirenb200_set_act_quant(self.handle, state.data_ptr(), 1, int(bool(training)),
                                                    float(averaging_const), int(qmin), int(qmax)))

    def set_fourier_encoding(self, B):
        """FourierNet: encoding.B [2, map_size / 2] (fourier.py:16-25)."""
        _lib.require_cuda(B, "B")
        self._enc_B = B.detach().to(torch.float32).contiguous()
        _lib.check(self.lib.sirenb200_set_fourier_encoding(self.handle, self._enc_B.data_ptr()))

    def workspace_bytes(self):
        return int(self.lib.sirenb200_workspace_bytes(self.handle))

    def profile(self, enable):
        _lib.check(self.lib.sirenb200_profile_enable(self.handle, int(bool(enable))))

    def profile_read(self):
        """{kind: (total_ms, launches)} of the tagged kernels since profile(True)."""
        n = len(_lib.PROFILE_KINDS)
        ms = (ctypes.c_float * n)()
        cnt = (ctypes.c_int32 * n)()
        _lib.check(self.lib.sirenb200_profile_read(self.handle, ms, cnt, n))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_lib.PROFILE_KINDS)}


def eval_metrics(pred, img):
    """Device tensor [mse, mse_8bit] (train_helper.py:48-57)."""
    _lib.require_cuda(pred, "pred")
    lib = _lib.load()
    out = torch.empty(2, dtype=torch.float32, device=pred.device)
    pred = pred.contiguous()
    img = img.contiguous()
    _lib.check(lib.sirenb200_eval_metrics(pred.data_ptr(), img.data_ptr(), pred.numel(), out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    return out


def apply_mask_(weight, mask):
    _lib.require_cuda(weight, "weight")
    lib = _lib.load()
    _lib.check(lib.sirenb200_apply_mask(weight.data_ptr(), mask.data_ptr(), weight.numel(),
                                        torch.cuda.current_stream().cuda_stream))
    return weight


def adam_step(params, grads, exp_avgs, exp_avg_sqs, masks, lr, beta1, beta2, eps, step, inv_scale=1.0,
              skip_flag=None, zero_grad=False):
    lib = _lib.load()
    n = len(params)
    numel = (ctypes.c_int64 * n)(*[p.numel() for p in params])
    mask_arr = _lib.ptr_array(masks) if masks is not None else None
    _lib.check(lib.sirenb200_adam_step(
        n, _lib.ptr_array(params), _lib.ptr_array(grads), _lib.ptr_array(exp_avgs),
        _lib.ptr_array(exp_avg_sqs), mask_arr, numel, float(lr), float(beta1), float(beta2), float(eps),
        int(step), float(inv_scale), None if skip_flag is None else skip_flag.data_ptr(),
        int(bool(zero_grad)), torch.cuda.current_stream().cuda_stream))


def kmeans_quantize(weight, bits, iter_limit=5, tol=1e-4, init_centers=None):
    """Returns (centroids[k], labels int64 like weight, new_weight) — kmeans.py:110-150."""
    _lib.require_cuda(weight, "weight")
    lib = _lib.load()
    w = weight.contiguous()
    cent = torch.empty(2 ** bits, dtype=torch.float32, device=w.device)
    ncent = torch.zeros(1, dtype=torch.int32, device=w.device)
    labels = torch.empty(w.shape, dtype=torch.int64, device=w.device)
    w_out = torch.empty_like(w)
    _lib.check(lib.sirenb200_kmeans_quantize(
        w.data_ptr(), w.numel(), int(bits), int(iter_limit), float(tol),
        None if init_centers is None else init_centers.contiguous().data_ptr(), cent.data_ptr(),
        ncent.data_ptr(), labels.data_ptr(), w_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    k = int(ncent.item())
    return cent[:k].clone(), labels, w_out


def fakequant_per_tensor(x, state, training=True, averaging_const=0.01, qmin=0, qmax=127, want_mask=False):
    """torch.fused_moving_avg_obs_fake_quant on a CUDA tensor: returns fake_quant(x) (and the straight-through
    mask); `state` ([4] fp32: running min, running max, scale, zero point) is updated in place."""
    _lib.require_cuda(x, "x")
    lib = _lib.load()
    x = x.contiguous()
    out = torch.empty_like(x)
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_mask else None
    _lib.check(lib.sirenb200_fakequant_per_tensor(
        x.data_ptr(), x.numel(), state.data_ptr(), int(bool(training)), float(averaging_const), int(qmin), int(qmax),
        out.data_ptr(), None if mask is None else mask.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return (out, mask) if want_mask else out


def fakequant_per_channel(weight, row_min=None, row_max=None, neg_div=128.0, pos_div=127.0):
    """Returns (codes int8, scales[out], dequantised weight) — per-output-row symmetric int8.  The default
    divisors are those of the installed torch's fused observer kernel; (127.5, 127.5) is torch 1.7."""
    _lib.require_cuda(weight, "weight")
    lib = _lib.load()
    w = weight.contiguous()
    rows, cols = w.shape
    codes = torch.empty(w.shape, dtype=torch.int8, device=w.device)
    scales = torch.empty(rows, dtype=torch.float32, device=w.device)
    w_out = torch.empty_like(w)
    have = row_min is not None and row_max is not None
    _lib.check(lib.sirenb200_fakequant_per_channel(
        w.data_ptr(), rows, cols, row_min.contiguous().data_ptr() if have else None,
        row_max.contiguous().data_ptr() if have else None, float(neg_div), float(pos_div),
        codes.data_ptr(), scales.data_ptr(), w_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return codes, scales, w_out
