"""Multi-step fit driver: the loop of compress.py:137-143 without a host synchronisation per step.

`train_epoch` keeps the reference's contract (returns this step's loss as a Python float, i.e. one
device->host sync per step).  `Fitter.steps(k)` runs k steps back to back on the stream — forward+MSE+
backward, (optional) gradient all-reduce for pixel-sharded fits, fused Adam (+mask), StepLR — and returns
the per-step losses as a DEVICE tensor; nothing blocks until the caller reads it.  Mask topology updates
(every `interval` steps) run host-side torch exactly where the reference runs them.
"""
import torch

from . import _lib
from .parallel import FlatGrads, shard_rows
from .utils.train_helper import FusedAdam


class Fitter:
    def __init__(self, model, optim, grid, img, lr_scheduler=None, mask=None, masking_cfg=None,
                 rank=0, world_size=1, group=None):
        if not isinstance(optim, FusedAdam):
            raise _lib.SirenB200Error("Fitter needs the FusedAdam from get_optimizer_lr_scheduler")
        _lib.require_cuda(grid, "grid")
        _lib.require_cuda(img, "img")
        self.model, self.optim, self.sched = model, optim, lr_scheduler
        self.mask, self.masking_cfg = mask, masking_cfg
        self.rank, self.world, self.group = rank, world_size, group
        H = int(grid.shape[0])
        self.row_begin, self.row_end = shard_rows(H, world_size, rank) if world_size > 1 else (0, H)
        self.grid = grid[self.row_begin:self.row_end].contiguous() if world_size > 1 else grid
        self.img = img[self.row_begin:self.row_end].contiguous() if world_size > 1 else img.contiguous()
        self.engine = model.engine_for(self.grid, self.row_begin, self.row_end, height=H)
        self.flat = FlatGrads(model.hot_parameters())
        self.flat.attach()
        self.inv_count = 1.0 / float(img.numel())
        self.step_index = 0

    def steps(self, k):
        """Run k fit steps; returns a device tensor [k] with each step's (pre-update) loss."""
        model, optim, flat = self.model, self.optim, self.flat
        model.train()
        losses = torch.empty(k, dtype=torch.float32, device=self.img.device)
        for i in range(k):
            model.run_weight_transforms()
            self.engine.forward_backward(model.kernel_parameters(), self.img, flat.views, flat.stats)
            if self.world > 1:
                flat.all_reduce(self.group)
                losses[i] = flat.stats[0] * self.inv_count
            else:
                losses[i:i + 1].copy_(flat.stats[1:2], non_blocking=True)
            optim.skip_flag = flat.stats[2:3]
            if self.mask:
                self.mask.step()
            else:
                optim.step()
            if self.sched:
                self.sched.step()
            if self.mask and self.masking_cfg is not None:
                if self.step_index <= self.masking_cfg["end_when"] and \
                        self.step_index % self.masking_cfg["interval"] == 0:
                    self.mask.update_connections()
            self.step_index += 1
        return losses
