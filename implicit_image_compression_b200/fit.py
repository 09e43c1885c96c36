"""Multi-step fit driver: the loop of compress.py:137-143 without a host synchronisation per step.

`train_epoch` keeps the reference's contract (returns this step's loss as a Python float, i.e. one
device->host sync per step).  `Fitter.steps(k)` runs k steps back to back on the stream — forward+MSE+
backward, (optional) gradient all-reduce for pixel-sharded fits, fused Adam (+mask), StepLR — and returns
the per-step losses as a DEVICE tensor; nothing blocks until the caller reads it.  Mask topology updates
(every `interval` steps) run host-side torch exactly where the reference runs them.

When nothing on the step needs the host (dense fit or dense-gradient masks, no quantisation transforms), ONE
step is captured in a CUDA graph and replayed: the learning-rate schedule, Adam's bias correction and the
loss history live in device memory (sirenb200_sched_step / sirenb200_adam_step_dev), so the ~16 kernel
launches of a step cost one graph launch.  This matters most for pixel-sharded fits, where a rank's kernels
are only a few microseconds long.
"""
import ctypes
import os

import torch

from . import _lib
from .parallel import FlatGrads, shard_rows
from .utils.train_helper import FusedAdam

_RING = 1 << 14


class Fitter:
    def __init__(self, model, optim, grid, img, lr_scheduler=None, mask=None, masking_cfg=None,
                 rank=0, world_size=1, group=None, use_graph=None):
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
        # sharded fits exchange gradients with the library's peer-memory kernel when every rank can map
        # every peer (one node); otherwise through torch.distributed
        self.peer_exchange = world_size > 1 and self.flat.use_peer_exchange(group)
        self.inv_count = 1.0 / float(img.numel())
        self.step_index = 0
        if use_graph is None:
            use_graph = os.environ.get("SIRENB200_GRAPH", "1") != "0"
        self.use_graph = bool(use_graph)
        self._graph = None
        self._graph_key = None
        self._warmed = False
        self.launches_per_step = 0  # library kernels per fit step (measured on the eager capture run)

    # ------------------------------------------------------------------ eager path
    def _eager_steps(self, k, losses, offset):
        model, optim, flat = self.model, self.optim, self.flat
        for i in range(k):
            model.run_weight_transforms()
            self.engine.forward_backward(model.kernel_parameters(), self.img, flat.views, flat.stats)
            for fn in model._post_backward:
                fn(model)
            if self.world > 1:
                flat.all_reduce(self.group)
                losses[offset + i] = flat.stats[0] * self.inv_count
            else:
                losses[offset + i:offset + i + 1].copy_(flat.stats[1:2], non_blocking=True)
            optim.skip_flag = flat.stats[2:3]
            if self.mask:
                self.mask.step()
            else:
                optim.step()
            if self.sched:
                self.sched.step()

    # ------------------------------------------------------------------ graph path
    def _graph_eligible(self):
        m, opt = self.model, self.optim
        if not self.use_graph or m._weight_transforms or m._post_backward or m._param_override:
            return False
        if self.mask is not None and not self.mask.dense_gradients:
            return False
        if len(opt.param_groups) != 1:
            return False
        steps_done = opt.param_groups[0].get("_fused_step", 0)
        if self.sched is not None:
            if type(self.sched).__name__ != "StepLR" or self.sched.last_epoch != steps_done:
                return False
        return True

    def _fit_args(self):
        """sirenb200_fit_t over the tables of _prepare_graph (kept alive in self._g)."""
        g, flat = self._g, self.flat
        if "fa" not in g:
            cast = lambda arr: ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))  # noqa: E731
            sharded = self.world > 1
            g["fa"] = _lib.FitArgs(
                n_tensors=g["n"], h_params=cast(g["p"]), h_grads=cast(g["g"]), h_exp_avg=cast(g["m"]),
                h_exp_avg_sq=cast(g["v"]), h_mask=cast(g["mask"]) if g["mask"] is not None else None,
                h_numel=g["numel"], beta1=g["beta1"], beta2=g["beta2"], eps=g["eps"],
                sched_state=g["state"].data_ptr(), stats=flat.stats.data_ptr(), loss_ring=g["ring"].data_ptr(),
                ring_len=_RING, loss_host=g["host_loss"].data_ptr(),
                comm=flat.comm.handle if (sharded and flat.comm is not None) else None,
                flat=flat.flat.data_ptr(), flat_n=flat.flat.numel(),
                inv_count=self.inv_count if sharded else 0.0)
        return g["fa"]

    def _graph_body(self):
        lib, flat, g = self.engine.lib, self.flat, self._g
        stream = torch.cuda.current_stream().cuda_stream
        if self.world == 1 or flat.comm is not None:
            # one C call: GEMM chain, then ONE kernel for partial reduction + peer exchange + loss + schedule + Adam
            _lib.check(lib.sirenb200_fit_step(self.engine.handle, self.img.data_ptr(),
                                              ctypes.byref(self._fit_args()), stream))
            self.engine.generation += 1
            return
        # torch.distributed exchange (a rank could not map a peer's memory): separate launches around NCCL
        self.engine.forward_backward(self.model.kernel_parameters(), self.img, flat.views, flat.stats)
        flat.all_reduce(self.group)
        _lib.check(lib.sirenb200_sched_step(g["state"].data_ptr(), flat.stats.data_ptr(), self.inv_count,
                                            g["ring"].data_ptr(), _RING, g["host_loss"].data_ptr(), stream))
        _lib.check(lib.sirenb200_adam_step_dev(
            g["n"], g["p"], g["g"], g["m"], g["v"], g["mask"], g["numel"], g["beta1"], g["beta2"],
            g["eps"], g["state"].data_ptr(), 1.0, flat.stats[2:3].data_ptr(), 0, stream))

    def _prepare_graph(self):
        opt = self.optim
        group = opt.param_groups[0]
        # the tensors the kernels train: the model's hot parameters, in the optimizer's order (frozen extras such as
        # FourierNet's encoding.B are in the optimizer but never get a gradient)
        hot = {id(p) for p in self.flat.params}
        params = [p for p in group["params"] if id(p) in hot]
        key = tuple(p.data_ptr() for p in params) + (self.img.data_ptr(),)
        if key == self._graph_key:
            return
        for p in params:
            opt._ensure_state(p)
        beta1, beta2 = group["betas"]
        dev = self.img.device
        masks = None
        mask_bufs = {}
        if self.mask is not None:
            for n, w in self.mask._masked_parameters():
                mask_bufs[w] = self.mask.mask_dict[n].clone()
            masks = _lib.ptr_array([mask_bufs.get(p) for p in params])
        n = len(params)
        # The device-side schedule (step counter, lr, bias corrections), the loss ring and its pinned host
        # mirror belong to the OPTIMIZER: every Fitter stepping it (e.g. train_epoch alternating between two
        # image buffers) shares them, so the device's step counter never has to be re-uploaded.
        shared = opt.__dict__.get("_sirenb200_sched")
        if shared is None or shared["state"].device != dev:
            shared = {
                "state": torch.zeros(8, dtype=torch.float64, device=dev),
                "ring": torch.zeros(_RING, dtype=torch.float32, device=dev),
                # pinned host float the schedule kernel also writes the loss to (unified addressing: the
                # device uses the host pointer), read by step_loss() after a stream synchronisation
                "host_loss": torch.zeros(2, dtype=torch.float32).pin_memory(),  # {loss, steps completed}
                "dev_state": None,  # what the device holds (skip the upload when unchanged)
            }
            opt.__dict__["_sirenb200_sched"] = shared
        self._shared = shared
        self._g = {
            "state": shared["state"], "ring": shared["ring"], "host_loss": shared["host_loss"],
            "n": n, "p": _lib.ptr_array([p.data for p in params]),
            "g": _lib.ptr_array(self._grad_views(params)),
            "m": _lib.ptr_array([opt.state[p]["exp_avg"] for p in params]),
            "v": _lib.ptr_array([opt.state[p]["exp_avg_sq"] for p in params]),
            "mask": masks, "mask_bufs": mask_bufs,
            "numel": (ctypes.c_int64 * n)(*[p.numel() for p in params]),
            "beta1": float(beta1), "beta2": float(beta2), "eps": float(group["eps"]),
        }
        self._sync_sched_state()
        # one eager run of the body (a real step) initialises everything lazily created, then capture
        self._graph = None
        self._graph_key = key

    def _grad_views(self, params):
        """The gradient tensors the captured step WRITES (this Fitter's flat views), in `params` order — not
        `p.grad`, which another Fitter on the same model may have rebound since."""
        by_param = {id(p): v for p, v in zip(self.flat.params, self.flat.views)}
        return [by_param[id(p)] for p in params]

    def _attach(self):
        flat = self.flat
        if any(p.grad is not v for p, v in zip(flat.params, flat.views)):
            flat.attach()

    def _sync_sched_state(self):
        group = self.optim.param_groups[0]
        steps_done = group.get("_fused_step", 0)
        if self.sched is not None:
            lr0, gamma, period = self.sched.base_lrs[0], self.sched.gamma, self.sched.step_size
        else:
            lr0, gamma, period = group["lr"], 1.0, 1 << 30
        beta1, beta2 = group["betas"]
        want = (float(steps_done), lr0, gamma, float(period), beta1, beta2)
        if want != self._shared["dev_state"]:
            self._g["state"].copy_(torch.tensor(want + (0.0, 0.0), dtype=torch.float64))
            self._shared["dev_state"] = want

    def _advance_host(self, k):
        group = self.optim.param_groups[0]
        group["_fused_step"] = group.get("_fused_step", 0) + k
        if self.sched is not None:
            self.sched.last_epoch += k
            lr = self.sched.base_lrs[0] * self.sched.gamma ** (self.sched.last_epoch // self.sched.step_size)
            group["lr"] = lr
            self.sched._last_lr = [lr]
        if self.mask is not None:
            for _ in range(k):  # host bookkeeping of Masking.step (core.py:690-702)
                decay = self.mask.prune_rate_decay
                if decay.mode == "cumulative":
                    decay.step(self.mask.mask_step, 1 - self.mask.stats.total_density)
                else:
                    decay.step(self.mask.mask_step)
                self.mask.mask_step += 1

    def _graph_steps(self, k, losses, offset):
        """k steps on the device-scheduled path; losses=None (k must be 1) returns that step's loss as a
        Python float instead of storing it (train_epoch's contract: one device->host read per step)."""
        self._prepare_graph()
        g = self._g
        if self.mask is not None:  # masks may have been replaced by update_connections()
            for n, w in self.mask._masked_parameters():
                g["mask_bufs"][w].copy_(self.mask.mask_dict[n])
        group = self.optim.param_groups[0]
        step0 = group.get("_fused_step", 0)
        done = 0
        self._sync_sched_state()
        if self._graph is None:
            if not self._warmed:
                n0 = _lib.launch_count()
                self._graph_body()  # eager: counts as a step
                self.launches_per_step = _lib.launch_count() - n0
                self._warmed = True
                done = 1
            if done < k:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._graph_body()
                # the capture itself does not execute anything
                self._graph = graph
        while done < k:
            self._graph.replay()
            done += 1
        value = None
        if losses is None:
            value = self._read_loss()
        elif k == 1:
            losses[offset:offset + 1] = g["ring"][step0 % _RING:step0 % _RING + 1]
        else:
            idx = (torch.arange(k, device=losses.device) + step0) % _RING
            losses[offset:offset + k] = g["ring"][idx]
        self._advance_host(k)
        ds = self._shared["dev_state"]
        if ds is not None:  # the device advanced its own step counter
            self._shared["dev_state"] = (ds[0] + k,) + ds[1:]
        return value

    def _read_loss(self):
        """The schedule kernel mirrors {loss, steps completed} into pinned host memory; after a stream
        synchronisation (which releases the GIL, unlike a Python-side poll of the count — measured: polling is
        not faster, the per-call gap is graph-launch latency) the loss is a plain host read."""
        torch.cuda.current_stream().synchronize()
        return float(self._shared["host_loss"][0])

    def step_loss(self):
        """One device-scheduled step, loss returned as a float; None when the step is not graph-eligible."""
        if self.masking_cfg is not None or not self._graph_eligible():
            return None
        if not self.model.training:
            self.model.train()
        self._attach()
        value = self._graph_steps(1, None, 0)
        self.step_index += 1
        return value

    def native_steps(self, k):
        """k dense steps through ONE C call (sirenb200_fit_steps: the whole loop lives in the library, nothing
        per step on the Python side; what a non-Python host would call).  Same device-side schedule as the
        graph path; returns the k losses as a device tensor.  No masks' topology updates, no transforms."""
        if not self._graph_eligible() or self.masking_cfg is not None:
            raise _lib.SirenB200Error("native_steps needs a graph-eligible (dense / dense-gradient) fit")
        if not self.model.training:
            self.model.train()
        self._attach()
        self._prepare_graph()
        g = self._g
        if self.mask is not None:
            for n, w in self.mask._masked_parameters():
                g["mask_bufs"][w].copy_(self.mask.mask_dict[n])
        self._sync_sched_state()
        step0 = self.optim.param_groups[0].get("_fused_step", 0)
        flat = self.flat
        fa = self._fit_args()
        if self.world > 1 and flat.comm is None:
            raise _lib.SirenB200Error("native_steps on a sharded fit needs the peer-memory exchange")
        k = min(int(k), _RING)
        _lib.check(self.engine.lib.sirenb200_fit_steps(self.engine.handle, k, self.img.data_ptr(),
                                                       ctypes.byref(fa),
                                                       torch.cuda.current_stream().cuda_stream))
        self._warmed = True
        idx = (torch.arange(k, device=self.img.device) + step0) % _RING
        losses = g["ring"][idx]
        self._advance_host(k)
        ds = self._shared["dev_state"]
        if ds is not None:
            self._shared["dev_state"] = (ds[0] + k,) + ds[1:]
        self.step_index += k
        return losses

    # ------------------------------------------------------------------ public
    def steps(self, k):
        """Run k fit steps; returns a device tensor [k] with each step's (pre-update) loss."""
        if not self.model.training:
            self.model.train()
        self._attach()
        losses = torch.empty(k, dtype=torch.float32, device=self.img.device)
        done = 0
        while done < k:
            # segment up to (and including) the next topology update
            seg = k - done
            update_after = False
            if self.mask and self.masking_cfg is not None:
                interval, end_when = self.masking_cfg["interval"], self.masking_cfg["end_when"]
                nxt = self.step_index + ((-self.step_index) % interval)  # next multiple of interval
                if nxt <= end_when and nxt - self.step_index + 1 <= seg:
                    seg = nxt - self.step_index + 1
                    update_after = True
            seg = min(seg, _RING)
            if self._graph_eligible():
                self._graph_steps(seg, losses, done)
            else:
                self._eager_steps(seg, losses, done)
            self.step_index += seg
            done += seg
            if update_after:
                self.mask.update_connections()
        return losses
