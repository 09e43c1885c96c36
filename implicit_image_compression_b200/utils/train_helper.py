"""Step functions of the fit (reference: implicit_image/utils/train_helper.py).

Same names, arguments and return values as the reference; the work is done by libsirenb200:
  train_epoch  -> one fused forward + MSE + backward (sirenb200_forward_backward) and one fused multi-tensor
                  Adam (+ mask apply) launch (sirenb200_adam_step)
  eval_epoch   -> fused forward + one metrics reduction
"""
import math
from contextlib import contextmanager

import torch
from torch.nn import functional as F

from .. import _lib
from .. import engine as _engine


@contextmanager
def _blank_context():
    yield


def get_device(device_str):
    """train_helper.py:62-66.  The B200 path needs CUDA; a CPU request is honoured as a device object but
    any compute on it raises (no CPU fallback)."""
    if device_str == "cuda" and torch.cuda.is_available():
        return torch.device(device_str)
    return torch.device("cpu")


class FusedAdam(torch.optim.Adam):
    """torch.optim.Adam whose step() is ONE multi-tensor kernel (sirenb200_adam_step), optionally fused with
    mask application.  State layout stays torch-compatible (`exp_avg`, `exp_avg_sq`, `step`) because
    Masking reads it (pipeline/masking/core.py:474-493, 631-650)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, **kwargs):
        unsupported = {k: v for k, v in kwargs.items() if k in ("weight_decay", "amsgrad") and v}
        if unsupported:
            raise _lib.SirenB200Error(f"FusedAdam: unsupported options {unsupported}")
        super().__init__(params, lr=lr, betas=betas, eps=eps)
        self.fused_masks = {}      # param -> mask tensor, applied inside the Adam kernel (set by Masking)
        self.skip_flag = None      # device float: non-zero -> skip the update (GradScaler semantics)
        self.inv_scale = 1.0       # gradients are multiplied by this inside the kernel (GradScaler.unscale_)
        self._step_count_fused = 0

    def _ensure_state(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["step"] = torch.tensor(0.0)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                _lib.require_cuda(p, "parameter")
            states = [self._ensure_state(p) for p in ps]
            group["_fused_step"] = group.get("_fused_step", 0) + 1
            step = group["_fused_step"]
            beta1, beta2 = group["betas"]
            masks = [self.fused_masks.get(p) for p in ps] if self.fused_masks else None
            for i in range(0, len(ps), 32):
                sl = slice(i, i + 32)
                _engine.adam_step([p.data for p in ps[sl]], [p.grad for p in ps[sl]],
                                  [s["exp_avg"] for s in states[sl]],
                                  [s["exp_avg_sq"] for s in states[sl]],
                                  masks[sl] if masks is not None else None, group["lr"], beta1, beta2,
                                  group["eps"], step, float(self.inv_scale), self.skip_flag, False)
        return loss

    def sync_state(self):
        """Materialise the torch-style per-parameter `step` tensors (cheap bookkeeping kept off the step)."""
        for group in self.param_groups:
            for p in group["params"]:
                if p in self.state and "exp_avg" in self.state[p]:
                    self.state[p]["step"] = torch.tensor(float(group.get("_fused_step", 0)))

    def state_dict(self):
        self.sync_state()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """A stock torch.optim.Adam checkpoint (or any state_dict without `_fused_step`) carries the step count
        only in the per-parameter `step` entries: seed the fused counter from them, otherwise bias correction
        would restart at t = 1 on warm moments.  The device-side schedule mirror is invalidated either way."""
        super().load_state_dict(state_dict)
        for group in self.param_groups:
            if "_fused_step" not in group:
                steps = [int(float(self.state[p]["step"])) for p in group["params"]
                         if p in self.state and "step" in self.state[p]]
                group["_fused_step"] = max(steps) if steps else 0
        shared = self.__dict__.get("_sirenb200_sched")
        if shared is not None:
            shared["dev_state"] = None


def get_optimizer_lr_scheduler(model, optim_cfg, quantize_mode=False):
    """train_helper.py:69-86: Adam(lr) + StepLR(2000, 0.5) (StepLR(1000, 0.5) in the quant phase).
    Shampoo (third-party torch_optimizer) is out of scope."""
    name = optim_cfg["name"] if isinstance(optim_cfg, dict) else optim_cfg.name
    kwargs = {k: v for k, v in dict(optim_cfg).items() if k != "name"}
    if name != "adam":
        raise _lib.SirenB200Error(f"optimizer '{name}' is not on the B200 hot path (only 'adam')")
    if "lr" in kwargs:
        kwargs["lr"] = float(kwargs["lr"])
    optim = FusedAdam(model.parameters(), **kwargs)
    period = 1000 if quantize_mode else 2000
    lr_scheduler = torch.optim.lr_scheduler.StepLR(optim, period, gamma=0.5)
    return optim, lr_scheduler


def setup_mask(model, optim, masking_cfg=None):
    """train_helper.py:89-129."""
    from ..pipeline.masking import Masking
    from ..pipeline.masking.funcs.decay import registry as decay_registry

    if masking_cfg and not masking_cfg["dense"]:
        if masking_cfg["decay_schedule"] == "magnitude-prune":
            kwargs = {"final_sparsity": 1 - masking_cfg["final_density"], "T_max": masking_cfg["end_when"],
                      "T_start": masking_cfg["start_when"], "interval": masking_cfg["interval"]}
        else:
            kwargs = {"prune_rate": masking_cfg["prune_rate"], "T_max": masking_cfg["end_when"]}
        decay = decay_registry[masking_cfg["decay_schedule"]](**kwargs)
        mask = Masking(optim, decay, input_size=(1, 1, 2), density=masking_cfg["density"],
                       dense_gradients=masking_cfg["dense_gradients"],
                       sparse_init=masking_cfg["sparse_init"], prune_mode=masking_cfg["prune_mode"],
                       growth_mode=masking_cfg["growth_mode"],
                       redistribution_mode=masking_cfg["redistribution_mode"])
        model.train()
        mask.add_module(model)
        return mask
    return None


def _fused_loss_and_grads(model, grid, img, loss_scale=1.0):
    """forward + F.mse_loss + backward in one library call; gradients (of loss * loss_scale) land in param.grad."""
    model.run_weight_transforms()
    params = model.hot_parameters()
    for p in params:
        if p.grad is None or p.grad.shape != p.shape:
            p.grad = torch.empty_like(p)
    eng = model.engine_for(grid)
    stats = eng.forward_backward(model.kernel_parameters(), img.contiguous(), [p.grad for p in params],
                                 loss_scale=loss_scale)
    for fn in model._post_backward:
        fn(model)
    return stats


def _graph_step(model, optim, grid, img, lr_scheduler, mask):
    """train_epoch through a cached fit.Fitter (CUDA-graph replay).  Returns None when the step is not
    eligible (quantisation transforms, sparse-gradient masks, foreign scheduler, ...)."""
    from ..fit import Fitter

    if model._weight_transforms or model._post_backward or model._param_override:
        return None
    cache = model.__dict__.setdefault("_step_fitters", {})
    key = (id(optim), id(lr_scheduler), id(mask), grid.data_ptr(), img.data_ptr(), tuple(img.shape))
    fitter = cache.get(key)
    if fitter is None:
        if len(cache) >= 4:
            cache.clear()
        fitter = Fitter(model, optim, grid, img, lr_scheduler, mask, None)
        cache[key] = fitter
    return fitter.step_loss()  # (re-)attaches param.grad to the views the captured step writes


def _scaler_update(scaler, found_inf, device):
    """GradScaler.update() with the library's non-finite flag as found_inf: the same device-side rule torch
    applies (torch._amp_update_scale_: x backoff on inf, x growth after growth_interval clean steps)."""
    if getattr(scaler, "_scale", None) is None:
        scaler._lazy_init_scale_growth_tracker(device)
    torch._amp_update_scale_(scaler._scale, scaler._growth_tracker, found_inf.reshape(1).to(torch.float32),
                             scaler._growth_factor, scaler._backoff_factor, scaler._growth_interval)


def train_epoch(model, optim, grid, img, **kwargs):
    """One fit step; returns this step's loss as a Python float (train_helper.py:132-185).

    kwargs: mask, pbar, lr_scheduler, scaler, criterion, context, preconditioner — as in the reference.
    The reference never enters autocast (it looks the context up under the wrong key), so a
    `scaler` means fp32 arithmetic with GradScaler around it (train_helper.py:156-175, core.py:679-685): the
    backward runs on loss * scale, the optimizer step is skipped when a scaled gradient is not finite, the
    gradients left in param.grad are unscaled, and the scale follows GradScaler.update()."""
    mask = kwargs.get("mask")
    pbar = kwargs.get("pbar")
    lr_scheduler = kwargs.get("lr_scheduler")
    criterion = kwargs.get("criterion", F.mse_loss)
    scaler = kwargs.get("scaler")
    use_scaler = scaler is not None and scaler.is_enabled()
    if kwargs.get("preconditioner"):
        raise _lib.SirenB200Error("preconditioners (EKFAC) are dead code in the reference and unsupported")

    if not model.training:
        model.train()
    fused = criterion is F.mse_loss and hasattr(model, "hot_parameters")
    if fused and isinstance(optim, FusedAdam) and not use_scaler:
        # one replay of the captured step graph (fit.Fitter) when nothing on the step needs the host;
        # the contract is unchanged: this step's loss as a float, param.grad populated, optimizer /
        # scheduler / mask bookkeeping advanced by one
        loss = _graph_step(model, optim, grid, img, lr_scheduler, mask)
        if loss is not None:
            if pbar:
                pbar.update(1)
            return loss
    if fused and use_scaler:
        if not isinstance(optim, FusedAdam):
            raise _lib.SirenB200Error("GradScaler on the fused path needs the FusedAdam optimizer")
        scale = float(scaler.get_scale())
        stats = _fused_loss_and_grads(model, grid, img, loss_scale=scale)
        host = stats.tolist()  # the step's one device->host read: [sum sq err, loss, non-finite flag, -]
        found_inf = host[2] != 0.0
        grads = [p.grad for p in model.hot_parameters()]
        torch._foreach_mul_(grads, 1.0 / scale)  # scaler.unscale_: what the reference leaves in param.grad
        optim.skip_flag, optim.inv_scale = None, 1.0
        if mask:
            mask.step(scaler, skip_optimizer=found_inf)
        elif not found_inf:
            optim.step()  # GradScaler.step() does not call optimizer.step() on a non-finite gradient
        _scaler_update(scaler, stats[2], stats.device)
        if pbar:
            pbar.update(1)
        if lr_scheduler:
            lr_scheduler.step()
        return host[1]
    if fused:
        stats = _fused_loss_and_grads(model, grid, img)
        loss_t = stats[1]
        if isinstance(optim, FusedAdam):
            optim.skip_flag = stats[2:3]
    else:
        optim.zero_grad()
        pred = model(grid)
        loss_t = criterion(pred, img)
        loss_t.backward()
        if isinstance(optim, FusedAdam):
            optim.skip_flag = None

    if mask:
        mask.step()
    else:
        optim.step()
    if pbar:
        pbar.update(1)
    if lr_scheduler:
        lr_scheduler.step()
    return loss_t.item()


@torch.no_grad()
def eval_epoch(model, grid, img, **kwargs):
    """(pred, loss, PSNR, PSNR_8bit) — train_helper.py:41-59.  Leaves the model in eval mode."""
    model.eval()
    pred = model(grid)
    m = _engine.eval_metrics(pred, img).tolist()
    mse, mse8 = m[0], m[1]
    psnr = 10 * math.log10(1 / mse) if mse > 0 else float("inf")
    psnr8 = 10 * math.log10(255 ** 2 / mse8) if mse8 > 0 else float("inf")
    return pred, mse, psnr, psnr8
