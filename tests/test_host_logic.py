"""CPU-only tests of the host-side logic of the product package (no GPU compute)."""
import os

import numpy as np
import pytest
import torch

from implicit_image_compression_b200.config import load_config
from implicit_image_compression_b200.models import Siren, registry
from implicit_image_compression_b200.parallel import FlatGrads, assign_replicas, shard_rows
from implicit_image_compression_b200.pipeline.masking.funcs import decay, init_scheme


def test_decay_schedules_match_reference(golden):
    g = golden("decay.npz")
    d = decay.registry["cosine"](prune_rate=0.1, T_max=1500)
    seq = [d.get_dr()]
    for s in range(60):
        d.step(s)
        seq.append(d.get_dr())
    np.testing.assert_array_equal(np.array(seq), g["cosine"])
    d = decay.registry["magnitude-prune"](final_sparsity=0.9, T_max=1500, T_start=5, interval=10)
    seq = []
    for s in range(200):
        d.step(s, 0.001 * s)
        seq.append(d.get_dr())
    np.testing.assert_array_equal(np.array(seq), g["magnitude_prune"])
    d = decay.registry["linear"](prune_rate=0.3, T_max=100)
    seq = []
    for s in range(120):
        d.step(s)
        seq.append(d.get_dr())
    np.testing.assert_array_equal(np.array(seq), g["linear"])


class _FakeMasking:
    def __init__(self, module, density):
        self.module, self.density = module, density
        self.mask_dict = {n: torch.zeros_like(p) for n, p in module.named_parameters() if "bias" not in n}
        self.baseline_nonzero = 0
        self.total_params = 0

    def remove_weight(self, name):
        self.mask_dict.pop(name, None)


@pytest.mark.parametrize("tag,scheme,density", [("rigl", "erdos-renyi-kernel", 0.5),
                                                ("snfs", "erdos-renyi-kernel", 0.3),
                                                ("pruning", "random", 1.0)])
def test_sparse_init_reproduces_reference_masks(golden, tag, scheme, density):
    """Same seed, same RNG consumption (including the reference's FLOP-probe draw) -> identical masks."""
    g = golden(f"masking_{tag}.npz")
    torch.manual_seed(0)
    model = Siren(depth=int(g["depth"]), hidden_size=int(g["hidden"]), first_omega_0=50, hidden_omega_0=30)
    for i, p in enumerate(model.parameters()):
        assert torch.equal(p.detach(), torch.from_numpy(g[f"init_param{i}"])) or tag != "pruning" or i > 0
    fm = _FakeMasking(model, density)
    torch.manual_seed(123)
    torch.rand(1, 1, 2)
    init_scheme.registry[scheme](fm)
    names = [str(n) for n in g["names"]]
    assert sorted(fm.mask_dict.keys()) == sorted(names)
    for n in names:
        assert torch.equal(fm.mask_dict[n], torch.from_numpy(g["init_mask/" + n])), n
    assert fm.baseline_nonzero == int(g["baseline_nonzero"])
    assert fm.total_params == int(g["total_params"])


def test_model_surface_matches_reference_naming():
    torch.manual_seed(0)
    m = registry["siren"](name="siren", depth=5, hidden_size=32, first_omega_0=50, hidden_omega_0=30,
                          outermost_linear=True, simulate_quantization=False, small_dense_density=0.25)
    names = [n for n, _ in m.named_parameters()]
    assert names == [f"layers.{i}.linear.{k}" for i in range(5) for k in ("weight", "bias")]
    assert m.layers[1].linear.weight.shape == (16, 16)  # int(32 * sqrt(0.25))
    assert all(isinstance(m.layers[i].linear, torch.nn.Linear) for i in range(5))
    assert sorted(m.state_dict().keys()) == sorted(names)
    import copy

    m2 = copy.deepcopy(m)
    assert all(torch.equal(a, b) for a, b in zip(m.parameters(), m2.parameters()))


def test_config_loader_honours_reference_keys():
    cfg = load_config(["mlp.hidden_size=256", "mlp.depth=6", "masking=Pruning", "quant=none",
                       "masking.final_density=0.1", "train.num_steps=10"])
    assert cfg.mlp.hidden_size == 256 and cfg.mlp.depth == 6 and cfg.mlp.first_omega_0 == 50
    assert cfg.masking.prune_mode == "global-magnitude" and cfg.masking.final_density == 0.1
    assert not cfg.quant and cfg.train.num_steps == 10 and cfg.optim.lr == 3e-4
    assert cfg.train.batch_width == cfg.img.width and cfg.exp_name.startswith("siren_")
    assert load_config(["mlp.width=64"]).mlp.hidden_size == 64
    d = load_config([])
    assert d.masking.name == "RigL" and d.quant.name == "KMeans" and d.quant.bits == 8


def test_row_sharding_and_replica_assignment():
    for H, world in ((2048, 8), (512, 3), (7, 4)):
        blocks = [shard_rows(H, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == H
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sizes = [e - b for b, e in blocks]
        assert max(sizes) - min(sizes) <= 1
    jobs = sorted(sum((assign_replicas(96, 8, r) for r in range(8)), []))
    assert jobs == list(range(96))


def test_flat_grads_views():
    ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))]
    fg = FlatGrads(ps)
    fg.attach()
    ps[0].grad.fill_(1.0)
    ps[1].grad.fill_(2.0)
    assert fg.flat[:6].eq(1).all() and fg.flat[6:11].eq(2).all() and fg.flat.numel() == 16  # 11 + 4 stats, padded to float4s
    fg.all_reduce()  # no process group: no-op


def _gloo_worker(rank, world, port, tmp):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import siren_oracle as O
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    H, W, depth, hidden = 12, 10, 3, 16
    params = O.siren_init(0, depth, hidden, 50.0, 30.0)
    grid, img = O.get_grid(H, W), O.synth_image(H, W, 0)
    b, e = shard_rows(H, world, rank)
    # per-rank partial gradients divided by the FULL element count (what the library emits per shard)
    loss_s, grads_s = O.siren_loss_and_grads(params, grid[b:e], img[b:e], 50.0, 30.0)
    frac = img[b:e].numel() / img.numel()
    holders = [torch.nn.Parameter(p.clone()) for p in params]
    fg = FlatGrads(holders)
    fg.attach()
    for v, gsh in zip(fg.views, grads_s):
        v.copy_(gsh * frac)
    fg.stats[0] = loss_s * img[b:e].numel()
    fg.all_reduce()
    loss_full, grads_full = O.siren_loss_and_grads(params, grid, img, 50.0, 30.0)
    ok = all((v - gf).norm() <= 1e-5 * gf.norm() + 1e-10 for v, gf in zip(fg.views, grads_full))
    ok = ok and abs(fg.stats[0].item() / img.numel() - loss_full.item()) < 1e-6
    torch.save(ok, os.path.join(tmp, f"ok{rank}.pt"))
    dist.destroy_process_group()


def test_pixel_sharded_gradients_sum_to_full_image_gloo(tmp_path):
    """world_size-2 gloo run of the N>1 host path: row shards + one all-reduce == full-image gradients."""
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(torch.load(tmp_path / f"ok{r}.pt") for r in range(2))


def test_sorted_probe_equals_torch_threshold_count():
    """global_magnitude_prune counts #(|w| > t) through one sort + upper_bound; torch compares an fp32 tensor
    with a Python float in fp32 (prune.py:74-76 of the reference relies on that), ties included."""
    import numpy as np
    import torch
    g = torch.Generator().manual_seed(3)
    w = torch.randn(4096, generator=g) * 0.05
    w[::7] = 0.0
    w[1::11] = w[5]  # duplicates
    mags = torch.sort(w.abs())[0].numpy()
    probes = [0.0, 1e-3, float(mags[100]), float(mags[100]) * (1 + 1e-9), float(mags[100]) * (1 - 1e-9),
              float(np.nextafter(mags[2000], np.float32(1))), 0.1, 1e9]
    for t in probes:
        direct = int((torch.abs(w) > t).sum())
        via_sort = mags.size - int(np.searchsorted(mags, np.float32(t), side="right"))
        assert direct == via_sort, t
