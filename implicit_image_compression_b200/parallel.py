"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL/NVLink).

The reference has no distributed code.  The fit shards naturally over pixels: the loss is a mean over
independent pixels and the parameters are tiny (<= 6.3 MB) and replicated.  Rank r owns a contiguous block
of image rows; each step it produces gradient partial sums already divided by the FULL image element count
(sirenb200_forward_backward), one all-reduce(sum) over a flat [grads | sum_sq_err, -, nonfinite, -] buffer
makes them the full-image gradients on every rank, and every rank applies the identical fused Adam, so the
weights stay bit-identical without a broadcast.  Sweeps (config c5) run one independent fit per GPU with no
communication ("replicas only").
"""
import torch


def shard_rows(height, world_size, rank):
    """Contiguous, balanced row blocks: the first (height % world) ranks get one extra row."""
    base, extra = divmod(height, world_size)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return begin, end


class FlatGrads:
    """One flat fp32 buffer holding every gradient plus 4 trailing stats floats; param.grad tensors are
    views into it so the all-reduce needs no packing kernels."""

    def __init__(self, params):
        self.params = list(params)
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total + 4, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(v)
            off += p.numel()
        self.stats = self.flat[total:total + 4]
        self.numel = total

    def attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce(self, group=None):
        """Sum gradients and stats over ranks (no-op without an initialised process group)."""
        if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(group) > 1:
            torch.distributed.all_reduce(self.flat, op=torch.distributed.ReduceOp.SUM, group=group)


def assign_replicas(num_jobs, world_size, rank):
    """Round-robin job indices of this rank for an independent-fit sweep."""
    return list(range(rank, num_jobs, world_size))
