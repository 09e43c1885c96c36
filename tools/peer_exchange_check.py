"""torchrun --nproc-per-node N tools/peer_exchange_check.py
Checks the peer-memory gradient exchange (sirenb200_comm_*) against NCCL on random data: equal to NCCL within
fp32 summation order, bit-identical across ranks, stable over many epochs and inside a CUDA graph; prints
the per-call time of both."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from implicit_image_compression_b200.parallel import PeerExchange  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 264707 + 4 + 1  # c2 parameter count + stats, padded
n = (n + 3) // 4 * 4
comm = PeerExchange.create(n)
g = torch.Generator(device=dev).manual_seed(100 + rank)
ok = True
for it in range(40):
    x = torch.randn(n, device=dev, generator=g)
    ref = x.clone()
    dist.all_reduce(ref)
    comm.all_reduce(x)
    torch.cuda.synchronize()
    err = (x - ref).abs().max().item()
    gathered = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(gathered, x)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if err > 1e-5 * world or not same:
        ok = False
        print(f"rank {rank} iter {it}: max err vs NCCL {err}, identical across ranks {same}")
# CUDA graph: the launch carries no per-step host state
x = torch.randn(n, device=dev, generator=g)
base = x.clone()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    comm.all_reduce(x)  # warm (epoch advance outside the graph is fine)
    torch.cuda.synchronize()
    x.copy_(base)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        comm.all_reduce(x)
    for _ in range(5):
        x.copy_(base)
        graph.replay()
torch.cuda.synchronize()
ref = base.clone()
dist.all_reduce(ref)
gerr = (x - ref).abs().max().item()
ok = ok and gerr <= 1e-5 * world


def clock(fn, k=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3


y = torch.randn(n, device=dev)
t_peer = clock(lambda: comm.all_reduce(y))
t_nccl = clock(lambda: dist.all_reduce(y))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"peer exchange check: ok={bool(flag.item())} graph_err={gerr:.2e} world={world} n={n} "
          f"peer_kernel_us={t_peer:.1f} nccl_us={t_nccl:.1f}")
torch.cuda.synchronize()
dist.barrier()
os._exit(0 if flag.item() else 1)
