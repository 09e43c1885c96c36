"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL/NVLink).

The reference has no distributed code.  The fit shards naturally over pixels: the loss is a mean over
independent pixels and the parameters are tiny (<= 6.3 MB) and replicated.  Rank r owns a contiguous block
of image rows; each step it produces gradient partial sums already divided by the FULL image element count
(sirenb200_forward_backward), one all-reduce(sum) over a flat [grads | sum_sq_err, -, nonfinite, -] buffer
makes them the full-image gradients on every rank, and every rank applies the identical fused Adam, so the
weights stay bit-identical without a broadcast.  Sweeps (config c5) run one independent fit per GPU with no
communication ("replicas only").
"""
import ctypes
import os

import torch

from . import _lib


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
        # padded to a multiple of 4 floats: the peer-memory exchange moves float4s
        self.flat = torch.zeros((total + 4 + 3) // 4 * 4, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(v)
            off += p.numel()
        self.stats = self.flat[total:total + 4]
        self.numel = total
        self.comm = None  # PeerExchange (one kernel over NVLink peer memory) or None (torch.distributed)

    def attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce(self, group=None):
        """Sum gradients and stats over ranks (no-op without an initialised process group)."""
        if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(group) > 1:
            if self.comm is not None:
                self.comm.all_reduce(self.flat)
            else:
                torch.distributed.all_reduce(self.flat, op=torch.distributed.ReduceOp.SUM, group=group)

    def use_peer_exchange(self, group=None):
        """Switch the exchange to the library's peer-memory kernel (CUDA tensors, one node, <= 8 ranks).
        Returns True when every rank could map every peer; otherwise all ranks keep torch.distributed."""
        dist = torch.distributed
        if not (self.flat.is_cuda and dist.is_available() and dist.is_initialized()):
            return False
        world = dist.get_world_size(group)
        if world < 2 or world > 8 or os.environ.get("SIRENB200_PEER_EXCHANGE", "1") == "0":
            return False
        comm = PeerExchange.create(self.flat.numel(), group, self.flat.device)
        if comm is not None:
            self.comm = comm
            return True
        return False


class PeerExchange:
    """sirenb200_comm_*: every rank allocates a peer-visible region, the 64-byte CUDA IPC handles travel
    through torch.distributed (any transport would do), and from then on one kernel per step sums the flat
    gradient buffer over NVLink (include/siren_b200.h)."""

    def __init__(self, lib, handle, rank, world):
        self.lib, self.handle, self.rank, self.world = lib, handle, rank, world

    @classmethod
    def create(cls, max_floats, group=None, device=None):
        """Collective constructor: EVERY rank runs the same sequence of collectives whatever fails locally
        (create -> MIN(ok) -> all_gather(handles) -> connect -> MIN(ok)); on any rank's failure all ranks
        destroy what they built and return None (the caller keeps torch.distributed)."""
        dist = torch.distributed
        lib = _lib.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device

        def all_ok(ok):
            flag = torch.tensor([1 if ok else 0], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            return int(flag.item()) == 1

        handle = ctypes.c_void_p()
        mine = ctypes.create_string_buffer(64)
        ok = lib.sirenb200_comm_create(rank, world, int(max_floats), ctypes.byref(handle)) == 0
        ok = ok and lib.sirenb200_comm_handle(handle, mine) == 0
        everyone = all_ok(ok)
        if everyone:
            handles = [None] * world
            dist.all_gather_object(handles, bytes(mine.raw), group=group)
            blob = ctypes.create_string_buffer(b"".join(handles), 64 * world)
            everyone = all_ok(lib.sirenb200_comm_connect(handle, blob) == 0)
        if not everyone:
            if handle:
                lib.sirenb200_comm_destroy(handle)
            return None
        return cls(lib, handle, rank, world)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def all_reduce(self, flat):
        _lib.check(self.lib.sirenb200_comm_allreduce(self.handle, flat.data_ptr(), flat.numel(),
                                                     torch.cuda.current_stream().cuda_stream))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.sirenb200_comm_destroy(self.handle)
            self.handle = ctypes.c_void_p()


def assign_replicas(num_jobs, world_size, rank):
    """Round-robin job indices of this rank for an independent-fit sweep."""
    return list(range(rank, num_jobs, world_size))
