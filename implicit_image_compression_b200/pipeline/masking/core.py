"""Sparse-training mask manager (reference: pipeline/masking/core.py).

Same public surface as the reference's `Masking` (add_module / step / update_connections / apply_mask /
mask_dict / stats / prune_rate ...).  What changed is where the work runs:
  * apply_mask        -> sirenb200_apply_mask (bit-exact 0/1 multiply, in place on the device)
  * step              -> ONE fused kernel: Adam + mask (train_helper.FusedAdam with `fused_masks`)
  * update_connections-> host-side torch on the device tensors every `interval` steps (not in the
                         per-step hot loop; SURVEY.md §8f ranks an on-device version as "next")
The FLOP counter (pipeline/masking/counting, logging only) is out of scope; the two `torch.rand(input_size)`
draws it made are kept so that mask initialisation consumes the CPU RNG stream exactly like the reference.
"""
import logging
import math
from dataclasses import dataclass, field

import numpy as np
import torch

from ... import _lib
from ... import engine as _engine
from .funcs.grow import registry as grow_registry
from .funcs.init_scheme import registry as init_registry
from .funcs.prune import registry as prune_registry
from .funcs.redistribute import registry as redistribute_registry


@dataclass
class LayerStats:
    """Layer-wise and global mask statistics (core.py:41-108)."""
    variance_dict: dict = field(default_factory=dict)
    zeros_dict: dict = field(default_factory=dict)
    nonzeros_dict: dict = field(default_factory=dict)
    removed_dict: dict = field(default_factory=dict)
    total_variance: float = 0
    total_zero: int = 0
    total_nonzero: int = 0
    total_removed: int = 0

    _KEYS = ("variance_dict", "zeros_dict", "nonzeros_dict", "removed_dict", "total_variance",
             "total_zero", "total_nonzero", "total_removed")

    def state_dict(self):
        return {k: getattr(self, k) for k in self._KEYS}

    def load_state_dict(self, *dicts, **kwargs):
        for d in list(dicts) + [kwargs]:
            for k, v in d.items():
                setattr(self, k, v)

    @property
    def total_density(self):
        total = self.total_zero + self.total_nonzero
        return self.total_nonzero / total if total else 0.0


class Masking:
    def __init__(self, optimizer, prune_rate_decay, density=0.2, sparse_init="random",
                 dense_gradients=False, prune_mode="magnitude", growth_mode="momentum",
                 redistribution_mode="momentum", prune_threshold=0.001, growth_threshold=0.001,
                 growth_increment=0.2, increment=0.2, tolerance=1e-6, input_size=(1, 3, 32, 32)):
        for mode, reg, what in ((sparse_init, init_registry, "sparse init"),
                                (growth_mode, grow_registry, "growth mode"),
                                (prune_mode, prune_registry, "prune mode"),
                                (redistribution_mode, redistribute_registry, "redistribution mode")):
            if mode not in reg:
                raise _lib.SirenB200Error(f"{what} '{mode}' is not available on the B200 path; "
                                          f"choose from {sorted(reg)}")
        self.optimizer = optimizer
        self.prune_rate_decay = prune_rate_decay
        self.density, self.sparse_init = density, sparse_init
        self.dense_gradients = dense_gradients
        self.prune_mode, self.growth_mode = prune_mode, growth_mode
        self.redistribution_mode = redistribution_mode
        self.prune_threshold, self.growth_threshold = prune_threshold, growth_threshold
        self.growth_increment, self.increment, self.tolerance = growth_increment, increment, tolerance
        self.input_size = input_size
        self.mask_dict = {}
        self.module = None
        self.mask_step = 0
        self.baseline_nonzero = 0
        self.total_params = 0
        self.adjusted_growth = 0
        self.adjustments = []
        self.name2prune_rate = {}
        self.stats = LayerStats()

    # ------------------------------------------------------------------ construction
    def add_module(self, module, lottery_mask_path=None):
        """core.py:220-248 + init() :387-423."""
        if lottery_mask_path is not None:
            raise _lib.SirenB200Error("lottery-ticket initialisation is out of scope")
        self.module = module
        torch.rand(*self.input_size)  # RNG parity: the reference draws this for its dense-FLOPs count
        for name, weight in module.named_parameters():
            self.mask_dict[name] = torch.zeros_like(weight, dtype=torch.float32, requires_grad=False)
        self.to_module_device_()
        self.remove_weight_partial_name("bias")
        init_registry[self.sparse_init](self)
        self.to_module_device_()
        self.apply_mask()
        self.stats.total_nonzero = self.baseline_nonzero
        self.stats.total_zero = self.total_params - self.baseline_nonzero
        torch.rand(*self.input_size)  # RNG parity: sparse-FLOPs count at init
        logging.info(f"Masking: {self.baseline_nonzero}/{self.total_params} weights active "
                     f"(target density {self.density})")

    def to_module_device_(self):
        for name, weight in self.module.named_parameters():
            if name in self.mask_dict:
                self.mask_dict[name] = self.mask_dict[name].to(weight.device)

    def remove_weight(self, name):
        if name in self.mask_dict:
            self.mask_dict.pop(name)
        elif name + ".weight" in self.mask_dict:
            self.mask_dict.pop(name + ".weight")
        else:
            logging.error(f"ERROR {name} not found.")

    def remove_weight_partial_name(self, partial_name):
        for name in list(self.mask_dict.keys()):
            if partial_name in name:
                self.mask_dict.pop(name)

    def remove_type(self, nn_type):
        for name, module in self.module.named_modules():
            if isinstance(module, nn_type):
                self.remove_weight(name)

    # ------------------------------------------------------------------ properties
    @property
    def prune_rate(self):
        return self.prune_rate_decay.get_dr()

    @property
    def prune_func(self):
        return prune_registry[self.prune_mode]

    @property
    def growth_func(self):
        return grow_registry[self.growth_mode]

    @property
    def redistribution_func(self):
        return redistribute_registry[self.redistribution_mode]

    @property
    def global_prune(self):
        return "global" in self.prune_mode

    def _masked_parameters(self):
        return [(n, w) for n, w in self.module.named_parameters() if n in self.mask_dict]

    # ------------------------------------------------------------------ mask application
    @torch.no_grad()
    def apply_mask(self):
        """core.py:272-279: w <- w * mask for every masked tensor (device kernel, in place)."""
        for name, weight in self._masked_parameters():
            _engine.apply_mask_(weight.data, self.mask_dict[name])

    @torch.no_grad()
    def apply_mask_gradients(self):
        """core.py:282-288."""
        for name, weight in self._masked_parameters():
            _engine.apply_mask_(weight.grad, self.mask_dict[name])

    @torch.no_grad()
    def reset_momentum(self):
        """core.py:631-650: mask Adam's moment buffers."""
        for name, weight in self._masked_parameters():
            st = self.optimizer.state[weight]
            if "exp_avg" in st:
                _engine.apply_mask_(st["exp_avg"], self.mask_dict[name])
                _engine.apply_mask_(st["exp_avg_sq"], self.mask_dict[name])
            elif "momentum_buffer" in st:
                _engine.apply_mask_(st["momentum_buffer"], self.mask_dict[name])

    def get_momentum_for_weight(self, weight):
        """core.py:474-493."""
        st = self.optimizer.state[weight]
        if "exp_avg" in st:
            return st["exp_avg"] / (torch.sqrt(st["exp_avg_sq"]) + 1e-08)
        if "momentum_buffer" in st:
            return st["momentum_buffer"]
        return []

    # ------------------------------------------------------------------ optimizer step
    def step(self, scaler=None, skip_optimizer=False):
        """core.py:671-702: optimizer step, then masks, then the prune-rate schedule.  GradScaler handling
        (scale, unscale, found-inf, update) is done by train_epoch around this call; `skip_optimizer` is its
        verdict for this step (scaler.step() skips optimizer.step() on a non-finite gradient, :679-682)."""
        opt = self.optimizer
        if skip_optimizer:
            self.apply_mask()
        elif hasattr(opt, "fused_masks"):
            opt.fused_masks = {w: self.mask_dict[n] for n, w in self._masked_parameters()}
            opt.step()  # Adam + mask multiply in one kernel
        else:
            opt.step()
            self.apply_mask()
        if not self.dense_gradients:
            self.reset_momentum()
        if self.prune_rate_decay.mode == "cumulative":
            self.prune_rate_decay.step(self.mask_step, 1 - self.stats.total_density)
        else:
            self.prune_rate_decay.step(self.mask_step)
        self.mask_step += 1

    # ------------------------------------------------------------------ topology update
    def gather_statistics(self):
        """core.py:425-464."""
        variance, nonzeros, zeros = {}, {}, {}
        total_variance, total_nonzero, total_zero = 0.0, 0, 0
        counts = []
        integral = True
        for name, weight in self._masked_parameters():
            mask = self.mask_dict[name]
            var = self.redistribution_func(self, name, weight, mask)
            integral = integral and not var.dtype.is_floating_point
            # (fp32 -> double is exact, counts are exact in a double: the values read back are those `.item()` gives)
            counts.append(torch.stack([var.double(), (mask == 1).sum().double(), (mask == 0).sum().double()]))
        if counts:
            counts = torch.stack(counts).tolist()  # one sync for all layers: statistic, active and inactive counts
        for (name, _), (var, nz, z) in zip(self._masked_parameters(), counts):
            variance[name] = int(var) if integral else var  # the `nonzero` rule counts (Python int in the reference)
            if not np.isnan(variance[name]):
                total_variance += variance[name]
            nonzeros[name], zeros[name] = int(nz), int(z)
            total_nonzero += int(nz)
            total_zero += int(z)
        assert total_variance, "Total variance is zero!"
        for name in variance:
            variance[name] /= total_variance
        self.stats = LayerStats(variance_dict=variance, nonzeros_dict=nonzeros, zeros_dict=zeros,
                                total_variance=total_variance, total_nonzero=total_nonzero,
                                total_zero=total_zero)

    def adjust_prune_rate(self):
        """core.py:250-269."""
        for name, mask in self.mask_dict.items():
            self.name2prune_rate[name] = self.prune_rate
            sparsity = self.stats.zeros_dict[name] / mask.numel()
            if sparsity < 0.2:
                expected_variance = 1.0 / len(self.stats.variance_dict.keys())
                actual_variance = self.stats.variance_dict[name]
                if expected_variance / actual_variance < 1.0:
                    self.name2prune_rate[name] = min(sparsity, self.name2prune_rate[name])

    def calc_redistributed_densities(self):
        """core.py:299-360 (the `global_magnitude` branch there can never trigger: the registry key is
        spelled `global-magnitude`, SURVEY.md App. A.6 — kept unreachable here as well)."""
        residual, mean_residual, name2regrowth, i = 9999, 0, {}, 0
        while residual > 0 and i < 1000:
            residual = 0
            for name in self.stats.variance_dict:
                max_regrowth = self.stats.zeros_dict[name] + self.stats.removed_dict[name]
                if name in name2regrowth:
                    regrowth = name2regrowth[name]
                else:
                    regrowth = round(self.stats.variance_dict[name]
                                     * (self.stats.total_removed + self.adjusted_growth))
                regrowth += mean_residual
                if regrowth > 0.99 * max_regrowth:
                    name2regrowth[name] = 0.99 * max_regrowth
                    residual += regrowth - name2regrowth[name]
                else:
                    name2regrowth[name] = regrowth
            mean_residual = residual / len(name2regrowth) if name2regrowth else 0
            i += 1
        return name2regrowth

    @torch.no_grad()
    def truncate_weights(self):
        """core.py:714-791: prune -> (redistribute) -> grow -> apply."""
        self.gather_statistics()
        self.adjust_prune_rate()
        total_nonzero_new = 0
        if self.global_prune:
            self.stats.total_removed = self.prune_func(self)
        else:
            kept = []
            for name, weight in self._masked_parameters():
                new_mask = self.prune_func(self, self.mask_dict[name], weight, name)
                kept.append(new_mask.sum())
                self.mask_dict[name] = new_mask
            kept = torch.stack(kept).tolist() if kept else []  # one sync for all layers
            for (name, _), k in zip(self._masked_parameters(), kept):
                removed = self.stats.nonzeros_dict[name] - int(k)
                self.stats.total_removed += removed
                self.stats.removed_dict[name] = removed
        if self.growth_mode == "none":
            total_nonzero_new = self.stats.total_nonzero - self.stats.total_removed
        else:
            redistribute = self.redistribution_mode not in ["nonzero", "none"]
            if redistribute:
                name2regrowth = self.calc_redistributed_densities()
            grown = []
            for name, weight in self._masked_parameters():
                num_growth = name2regrowth[name] if redistribute else self.stats.removed_dict[name]
                new_mask = self.growth_func(self, name, num_growth, weight)
                grown.append(new_mask.sum())
                self.mask_dict.pop(name)
                self.mask_dict[name] = new_mask.float()
            total_nonzero_new += sum(torch.stack(grown).tolist()) if grown else 0  # one sync for all layers
        self.apply_mask()
        if not self.dense_gradients:
            self.reset_momentum()
            self.apply_mask_gradients()
        self.mask_step += 1
        self.adjustments.append(self.baseline_nonzero - total_nonzero_new)
        self.adjusted_growth = (0.25 * self.adjusted_growth + (0.75 * self.adjustments[-1])
                                + np.mean(self.adjustments))
        self.gather_statistics()

    def update_connections(self):
        """core.py:793-801."""
        self.truncate_weights()

    # ------------------------------------------------------------------ (de)serialisation
    def state_dict(self):
        return {"baseline_nonzero": self.baseline_nonzero, "masks": self.mask_dict,
                "stats": self.stats.state_dict(), "mask_step": self.mask_step,
                "total_params": self.total_params}

    def load_state_dict(self, *dicts, **kwargs):
        for d in list(dicts) + [kwargs]:
            for k, v in d.items():
                if k == "stats":
                    self.stats.load_state_dict(v)
                elif k == "masks":
                    self.mask_dict = v
                else:
                    setattr(self, k, v)
