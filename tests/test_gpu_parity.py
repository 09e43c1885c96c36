"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the
oracle and the committed golden vectors.

Tolerances (stated per SURVEY.md §7.2 item 3 / BASELINE.json north_star):
  fp32 path   : forward <= 2e-6 abs, gradients <= 5e-6 relative (fp32 summation order only)
  f16tc path  : fp16 operands (10-bit mantissa, TF32-class) with fp32 accumulation; stashed activations
                carry a 9-bit mantissa + cos-sign bit.  forward <= 5e-4 abs on pred, gradients <= 1e-2
                relative (L2) per tensor, loss <= 1e-4 relative.
  masks, k-means codes, int8 codes : bit-exact.
"""
import copy
import math

import numpy as np
import pytest
import torch

import siren_oracle as O

pytestmark = pytest.mark.gpu

TOL = {"fp32": dict(pred=2e-6, grad=5e-6, loss=1e-6), "f16tc": dict(pred=5e-4, grad=1e-2, loss=1e-4)}


def _pkg():
    from implicit_image_compression_b200 import engine
    from implicit_image_compression_b200.data import get_grid, synth_image
    from implicit_image_compression_b200.models import Siren
    from implicit_image_compression_b200.utils import train_helper

    return engine, get_grid, synth_image, Siren, train_helper


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _model_from_golden(g, precision):
    _, _, _, Siren, _ = _pkg()
    torch.manual_seed(0)
    model = Siren(depth=int(g["depth"]), hidden_size=int(g["hidden"]), first_omega_0=50, hidden_omega_0=30,
                  precision=precision)
    for i, p in enumerate(model.parameters()):
        assert torch.equal(p.detach(), torch.from_numpy(g[f"param{i}"])), "init RNG stream differs"
    return model.cuda()


CASES = [("fit_d3_w16.npz", "fp32"), ("fit_d4_w128.npz", "fp32"), ("fit_d4_w128.npz", "f16tc"),
         ("fit_d3_w256.npz", "fp32"), ("fit_d3_w256.npz", "f16tc"), ("fit_d3_w64_c1small.npz", "fp32")]


@pytest.mark.parametrize("name,precision", CASES)
def test_forward_loss_grads_vs_reference_golden(golden, name, precision):
    g = golden(name)
    tol = TOL[precision]
    model = _model_from_golden(g, precision)
    grid = torch.from_numpy(g["grid"]).cuda()
    img = torch.from_numpy(g["img"]).cuda()
    with torch.no_grad():
        pred = model(grid)
    assert (pred.cpu() - torch.from_numpy(g["pred"])).abs().max().item() <= tol["pred"]
    params = model.hot_parameters()
    grads = [torch.full_like(p, float("nan")) for p in params]
    stats = model.engine_for(grid).forward_backward(model.kernel_parameters(), img, grads).tolist()
    assert abs(stats[1] - float(g["loss"])) <= tol["loss"] * float(g["loss"])
    assert stats[2] == 0.0
    for i, gr in enumerate(grads):
        assert _rel(gr, torch.from_numpy(g[f"grad{i}"])) <= tol["grad"], f"grad {i}"


@pytest.mark.parametrize("name,precision,rtol", [("fit_d3_w16.npz", "fp32", 2e-4),
                                                 ("fit_d3_w64_c1small.npz", "fp32", 1e-3),
                                                 ("fit_d4_w128.npz", "fp32", 5e-4),
                                                 ("fit_d4_w128.npz", "f16tc", 5e-3)])
def test_train_epoch_trajectory_vs_reference(golden, name, precision, rtol):
    """Drop-in train_epoch (fused fwd/bwd + fused Adam + StepLR) tracks the reference's loss curve."""
    _, _, _, _, th = _pkg()
    g = golden(name)
    model = _model_from_golden(g, precision)
    grid, img = torch.from_numpy(g["grid"]).cuda(), torch.from_numpy(g["img"]).cuda()
    optim, sched = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
    losses = [th.train_epoch(model, optim, grid, img, lr_scheduler=sched) for _ in range(len(g["losses"]))]
    np.testing.assert_allclose(losses, g["losses"], rtol=rtol)
    _, loss, psnr, psnr8 = th.eval_epoch(model, grid, img)
    assert abs(psnr - g["eval"][1]) <= (0.02 if precision == "fp32" else 0.1)  # dB
    assert abs(psnr8 - g["eval"][2]) <= 0.1
    if f"param_after0" in g.files and precision == "fp32":
        for i, p in enumerate(model.parameters()):
            ref = torch.from_numpy(g[f"param_after{i}"])
            assert (p.detach().cpu() - ref).abs().max() <= 5e-5 * (1 + ref.abs().max())


@pytest.mark.parametrize("precision,hidden", [("fp32", 48), ("f16tc", 128)])
def test_edge_shapes_and_arbitrary_grids(precision, hidden):
    """1x1 and 5x5 grids (the reference's FLOP probe / smoke inputs), ragged pixel counts (not a multiple
    of the 128-row tile) and non-separable coordinate tensors."""
    _, get_grid, _, Siren, _ = _pkg()
    torch.manual_seed(5)
    model = Siren(depth=4, hidden_size=hidden, first_omega_0=50, hidden_omega_0=30, precision=precision)
    ref = [p.detach().clone() for p in model.parameters()]
    model = model.cuda()
    tol = TOL[precision]["pred"]
    for shape in ((1, 1), (5, 5), (10, 10), (13, 29)):
        grid = torch.rand(*shape, 2)
        with torch.no_grad():
            pred = model(grid.cuda())
        want = O.siren_forward(ref, grid, 50.0, 30.0)
        assert pred.shape == want.shape
        assert (pred.cpu() - want).abs().max().item() <= tol, shape
    grid = get_grid(7, 300)
    with torch.no_grad():
        pred = model(grid.cuda())
    assert (pred.cpu() - O.siren_forward(ref, grid, 50.0, 30.0)).abs().max().item() <= tol


@pytest.mark.parametrize("depth,H,W", [(3, 24, 40), (4, 33, 47)])
def test_hidden_512_tensor_core_path_vs_oracle(depth, H, W):
    """hidden 512 (configs 3 and 5): streamed-B rowgemm with two output parts, 4x2 colgemm blocks."""
    _, _, _, Siren, _ = _pkg()
    torch.manual_seed(0)
    model = Siren(depth=depth, hidden_size=512, first_omega_0=50, hidden_omega_0=30, precision="f16tc")
    ref = O.siren_init(0, depth, 512, 50.0, 30.0)
    for p, r in zip(model.parameters(), ref):
        assert torch.equal(p.detach(), r)
    model = model.cuda()
    grid, img = O.get_grid(H, W), O.synth_image(H, W, 0)
    loss, grads = O.siren_loss_and_grads(ref, grid, img, 50.0, 30.0)
    with torch.no_grad():
        pred = model(grid.cuda())
    assert (pred.cpu() - O.siren_forward(ref, grid, 50.0, 30.0)).abs().max().item() <= TOL["f16tc"]["pred"]
    out = [torch.full_like(p, float("nan")) for p in model.hot_parameters()]
    stats = model.engine_for(grid.cuda()).forward_backward(model.kernel_parameters(), img.cuda(), out)
    assert abs(stats[1].item() - loss.item()) <= TOL["f16tc"]["loss"] * loss.item()
    for i, (a, b) in enumerate(zip(out, grads)):
        assert _rel(a, b) <= TOL["f16tc"]["grad"], f"grad {i}"


def test_sine_output_layer_variant():
    _, get_grid, synth_image, Siren, _ = _pkg()
    torch.manual_seed(2)
    for precision, hidden in (("fp32", 24), ("f16tc", 128)):
        model = Siren(depth=3, hidden_size=hidden, first_omega_0=50, hidden_omega_0=30,
                      outermost_linear=False, precision=precision)
        ref = [p.detach().clone() for p in model.parameters()]
        model = model.cuda()
        grid, img = O.get_grid(9, 20), O.synth_image(9, 20, 0)
        loss, grads = O.siren_loss_and_grads(ref, grid, img, 50.0, 30.0, outermost_linear=False)
        out = [torch.empty_like(p) for p in model.hot_parameters()]
        stats = model.engine_for(grid.cuda()).forward_backward(model.kernel_parameters(), img.cuda(), out)
        assert abs(stats[1].item() - loss.item()) <= 2e-4 * loss.item()
        for a, b in zip(out, grads):
            assert _rel(a, b) <= (1e-5 if precision == "fp32" else 2e-2)


@pytest.mark.parametrize("precision,hidden", [("fp32", 32), ("f16tc", 128)])
def test_autograd_bridge_custom_criterion(precision, hidden):
    """model(grid) participates in autograd (custom criterion path of train_epoch)."""
    _, _, _, Siren, th = _pkg()
    torch.manual_seed(0)
    model = Siren(depth=3, hidden_size=hidden, precision=precision, first_omega_0=50, hidden_omega_0=30)
    ref = [p.detach().clone().requires_grad_(True) for p in model.parameters()]
    model = model.cuda()
    grid, img = O.get_grid(11, 17), O.synth_image(11, 17, 3)
    crit = lambda a, b: (a - b).abs().mean()  # noqa: E731
    pred = O.siren_forward(ref, grid, 50.0, 30.0)
    crit(pred, img).backward()
    optim, _ = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
    loss = th.train_epoch(model, optim, grid.cuda(), img.cuda(), criterion=crit)
    assert abs(loss - crit(pred, img).item()) <= 1e-4
    # after the step p.grad still holds this step's gradients
    for p, r in zip(model.parameters(), ref):
        assert _rel(p.grad, r.grad) <= (1e-5 if precision == "fp32" else 3e-2)


def test_fused_adam_matches_oracle_and_torch():
    engine, *_ = _pkg()
    torch.manual_seed(1)
    shapes = [(64, 2), (64,), (64, 64), (64,), (3, 64), (3,)]
    p0 = [torch.randn(s) * 0.05 for s in shapes]
    gs = [[torch.randn(s) * 1e-3 for s in shapes] for _ in range(5)]
    mask = (torch.rand(64, 64) < 0.5).float()
    pr, mr, vr = [p.clone() for p in p0], [torch.zeros_like(p) for p in p0], [torch.zeros_like(p) for p in p0]
    pc = [p.clone().cuda() for p in p0]
    mc = [torch.zeros_like(p) for p in pc]
    vc = [torch.zeros_like(p) for p in pc]
    masks = [None, None, mask.cuda(), None, None, None]
    for step, g in enumerate(gs, 1):
        lr = O.steplr(3e-4, step - 1, period=2)
        for i in range(len(pr)):
            pr[i], mr[i], vr[i] = O.adam_step(pr[i], g[i], mr[i], vr[i], step, lr)
        pr[2] = O.apply_mask(pr[2], mask)
        engine.adam_step(pc, [x.cuda() for x in g], mc, vc, masks, lr, 0.9, 0.999, 1e-8, step)
    for a, b in zip(pc, pr):
        assert (a.cpu() - b).abs().max().item() <= 2e-7
    assert torch.equal((pc[2] == 0).cpu() | (mask == 1), torch.ones(64, 64, dtype=torch.bool))
    # skip flag: nothing changes except the mask multiply
    before = [x.clone() for x in pc]
    flag = torch.ones(1, device="cuda")
    engine.adam_step(pc, [x.cuda() for x in gs[0]], mc, vc, masks, 3e-4, 0.9, 0.999, 1e-8, 6, skip_flag=flag)
    assert all(torch.equal(a, b) for a, b in zip(pc, before))


def test_mask_apply_bit_exact_and_metrics():
    engine, *_ = _pkg()
    torch.manual_seed(0)
    w = torch.randn(300, 257)
    mk = (torch.rand(300, 257) < 0.1).float()
    wc = w.cuda()
    engine.apply_mask_(wc, mk.cuda())
    assert torch.equal(wc.cpu(), O.apply_mask(w, mk))
    a, b = torch.rand(64, 96, 3), torch.rand(64, 96, 3)
    mse, psnr, psnr8 = O.eval_metrics(a, b)
    got = engine.eval_metrics(a.cuda(), b.cuda()).tolist()
    assert abs(got[0] - mse) <= 1e-6 * mse
    assert abs(10 * math.log10(255 ** 2 / got[1]) - psnr8) <= 1e-4


def test_kmeans_codes_bit_exact_vs_reference(golden):
    engine, *_ = _pkg()
    g = golden("quant.npz")
    w = torch.from_numpy(g["w"]).cuda()
    for bits in (4, 8):
        init = torch.from_numpy(g[f"kmeans{bits}_init"]).cuda()
        c, l, nw = engine.kmeans_quantize(w, bits, init_centers=init)
        assert torch.equal(c.cpu(), torch.from_numpy(g[f"kmeans{bits}_centroids"]))
        assert torch.equal(l.cpu(), torch.from_numpy(g[f"kmeans{bits}_labels"]))
        assert torch.equal(nw.cpu(), torch.from_numpy(g[f"kmeans{bits}_weight"]))
    # in-kernel linspace == torch.linspace on the device (what the reference calls on a CUDA model)
    nz = w[w != 0]
    init = torch.linspace(nz.min().item(), nz.max().item(), 255, device="cuda")
    a = engine.kmeans_quantize(w, 8, init_centers=init)
    b = engine.kmeans_quantize(w, 8)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    # larger, pruned layer against the oracle
    torch.manual_seed(3)
    big = (torch.rand(256, 256) - 0.5) * 0.03
    big[torch.rand(256, 256) < 0.9] = 0
    co, lo, wo = O.kmeans_quantize(big, 8)
    nzb = big[big != 0]
    init = torch.linspace(nzb.min().item(), nzb.max().item(), 255)
    cg, lg, wg = engine.kmeans_quantize(big.cuda(), 8, init_centers=init.cuda())
    assert torch.equal(cg.cpu(), co) and torch.equal(lg.cpu(), lo) and torch.equal(wg.cpu(), wo)


def test_int8_fake_quant_codes_bit_exact(golden):
    engine, *_ = _pkg()
    g = golden("quant.npz")
    q, s, d = engine.fakequant_per_channel(torch.from_numpy(g["fq_w"]).cuda())
    assert np.array_equal(q.cpu().numpy(), g["fq_codes"])
    assert torch.equal(s.cpu(), torch.from_numpy(g["fq_scale"]))
    assert torch.equal(d.cpu(), torch.from_numpy(g["fq_deq"]))
    w = torch.randn(16, 33) * 0.02
    qo, so, do = O.fake_quant_per_channel_weight(w, neg_div=127.5, pos_div=127.5)
    q, s, d = engine.fakequant_per_channel(w.cuda(), neg_div=127.5, pos_div=127.5)
    assert torch.equal(q.cpu(), qo) and torch.equal(s.cpu(), so) and torch.equal(d.cpu(), do)


@pytest.mark.parametrize("tag", ["pruning", "rigl", "snfs"])
def test_masking_update_replays_reference_bit_exact(golden, tag):
    """Load the reference's state before each update_connections() call and replay the update with the
    product Masking on the GPU: masks and weights afterwards must be bit-identical."""
    _, _, _, Siren, th = _pkg()
    from implicit_image_compression_b200.pipeline.masking import Masking
    from implicit_image_compression_b200.pipeline.masking.funcs.decay import registry as decay_registry

    g = golden(f"masking_{tag}.npz")
    cfgs = {
        "pruning": dict(density=1.0, sparse_init="random", dense_gradients=True, growth_mode="none",
                        prune_mode="global-magnitude", redistribution_mode="none"),
        "rigl": dict(density=0.5, sparse_init="erdos-renyi-kernel", dense_gradients=True,
                     growth_mode="absolute-gradient", prune_mode="magnitude", redistribution_mode="none"),
        "snfs": dict(density=0.3, sparse_init="erdos-renyi-kernel", dense_gradients=True,
                     growth_mode="momentum", prune_mode="magnitude", redistribution_mode="momentum"),
    }[tag]
    names = [str(n) for n in g["names"]]
    torch.manual_seed(0)
    model = Siren(depth=int(g["depth"]), hidden_size=int(g["hidden"]), first_omega_0=50, hidden_omega_0=30,
                  precision="fp32").cuda()
    optim, _ = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})

    class _FixedRate:
        mode = "current"

        def __init__(self):
            self.rate = 0.0

        def get_dr(self):
            return self.rate

        def step(self, *a, **k):
            pass

    decay = _FixedRate()
    torch.manual_seed(123)
    mask = Masking(optim, decay, input_size=(1, 1, 2), **cfgs)
    mask.add_module(model)
    for n in names:  # same seed + same RNG consumption -> the reference's initial masks
        assert torch.equal(mask.mask_dict[n].cpu(), torch.from_numpy(g["init_mask/" + n])), n
    assert mask.baseline_nonzero == int(g["baseline_nonzero"])
    pmap = dict(model.named_parameters())
    for u in range(int(g["num_updates"])):
        pre = f"upd{u}/"
        for n, p in pmap.items():
            p.data = torch.from_numpy(g[pre + "w_before/" + n]).cuda()
            p.grad = torch.from_numpy(g[pre + "g_before/" + n]).cuda()
            st = optim.state[p]
            st["exp_avg"] = torch.from_numpy(g[pre + "m_before/" + n]).cuda()
            st["exp_avg_sq"] = torch.from_numpy(g[pre + "v_before/" + n]).cuda()
        for n in names:
            mask.mask_dict[n] = torch.from_numpy(g[pre + "mask_before/" + n]).cuda()
        sb = g[pre + "scalars_before"]
        mask.prune_threshold, decay.rate = float(sb[0]), float(sb[1])
        mask.mask_step, mask.adjusted_growth = int(sb[2]), float(sb[3])
        mask.stats.total_nonzero, mask.stats.total_zero = int(sb[4]), int(sb[5])
        mask.adjustments = list(g[pre + "adjustments_before"])
        mask.update_connections()
        sa = g[pre + "scalars_after"]
        for n in names:
            assert torch.equal(mask.mask_dict[n].cpu(), torch.from_numpy(g[pre + "mask_after/" + n])), (u, n)
        for n, p in pmap.items():
            assert torch.equal(p.detach().cpu(), torch.from_numpy(g[pre + "w_after/" + n])), (u, n)
        assert mask.prune_threshold == float(sa[0])
        assert mask.stats.total_nonzero == int(sa[4]) and mask.stats.total_zero == int(sa[5])
        assert abs(mask.adjusted_growth - float(sa[3])) < 1e-9


def test_masked_training_keeps_pruned_weights_at_zero():
    _, get_grid, synth_image, Siren, th = _pkg()
    torch.manual_seed(0)
    model = Siren(depth=4, hidden_size=128, first_omega_0=50, hidden_omega_0=30).cuda()
    grid, img = get_grid(32, 48, "cuda"), synth_image(32, 48, 0, device="cuda")
    optim, sched = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
    cfg = dict(name="RigL", density=0.5, sparse_init="erdos-renyi-kernel", dense_gradients=True,
               growth_mode="absolute-gradient", prune_mode="magnitude", redistribution_mode="none",
               dense=False, prune_rate=0.1, decay_schedule="cosine", end_when=20, interval=4)
    mask = th.setup_mask(model, optim, cfg)
    first = None
    for i in range(12):
        loss = th.train_epoch(model, optim, grid, img, lr_scheduler=sched, mask=mask)
        first = loss if first is None else first
        if i % cfg["interval"] == 0:
            mask.update_connections()
        for n, p in model.named_parameters():
            if n in mask.mask_dict:
                assert torch.equal(p.detach() * mask.mask_dict[n], p.detach()), n
    assert loss < first
    assert abs(mask.stats.total_density - 0.5) < 0.02


def test_quantize_context_kmeans_end_to_end(golden):
    """Quantize(KMeans) drop-in: hooks fire at weight load for train and eval forwards; after convert()
    the weights are exactly centroids[labels] with <= 2^bits distinct values."""
    _, _, _, Siren, th = _pkg()
    from implicit_image_compression_b200.pipeline.quant import Quantize

    g = golden("quant.npz")
    torch.manual_seed(0)
    model = Siren(depth=4, hidden_size=32, first_omega_0=50, hidden_omega_0=30, precision="fp32")
    for i, p in enumerate(model.parameters()):
        assert torch.equal(p.detach(), torch.from_numpy(g[f"q_param{i}"]))
    model = model.cuda()
    grid, img = O.get_grid(12, 12).cuda(), O.synth_image(12, 12, 2).cuda()
    optim, sched = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4}, quantize_mode=True)
    qcfg = dict(name="KMeans", bits=4, skip_ll=["layers.0.linear", "layers.3.linear"], num_steps=3)
    with Quantize(model, optim, qcfg) as q:
        losses = [th.train_epoch(model, optim, grid, img, lr_scheduler=sched) for _ in range(3)]
    qm = q.convert()
    np.testing.assert_allclose(losses, g["q_losses"], rtol=2e-3)
    for name in ("layers.1.linear", "layers.2.linear"):
        mod = dict(qm.named_modules())[name]
        lab = mod.labeled_weight.cpu()
        ref_lab = torch.from_numpy(g[f"q_labels/{name}"])
        assert (lab == ref_lab).float().mean().item() >= 0.99
        assert mod.centroids.numel() <= 16
        assert torch.equal(mod.weight.data, mod.centroids[mod.labeled_weight])
        # includes the reference's code-book SGD step after the last backward (kmeans.py:170-177)
        np.testing.assert_allclose(mod.centroids.cpu().numpy(), g[f"q_centroids/{name}"], rtol=2e-3,
                                   atol=2e-7)
    assert not hasattr(dict(qm.named_modules())["layers.0.linear"], "labeled_weight")


def test_quantize_context_qat_weights_only():
    _, get_grid, synth_image, Siren, th = _pkg()
    from implicit_image_compression_b200.pipeline.quant import Quantize

    torch.manual_seed(0)
    model = Siren(depth=4, hidden_size=128, first_omega_0=50, hidden_omega_0=30).cuda()
    grid, img = get_grid(24, 32, "cuda"), synth_image(24, 32, 0, device="cuda")
    optim, sched = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4}, quantize_mode=True)
    master0 = [p.detach().clone() for p in model.parameters()]
    # activations=False: the weights-only variant, which stays on the tensor-core path (the full QAT of the
    # reference — weights AND activation observers — is tests/test_gpu_qat.py)
    with Quantize(model, optim, dict(name="QAT", qconfig="fbgemm", activations=False)) as q:
        l0 = th.train_epoch(model, optim, grid, img, lr_scheduler=sched)
        # gradients were taken at the fake-quantised weights and applied to the fp32 master weights
        wq = O.fake_quant_per_channel_weight(master0[2].cpu())[2]
        ref = [m.cpu() for m in master0]
        ref[2] = wq
        for i in (0, 4, 6):
            ref[i] = O.fake_quant_per_channel_weight(master0[i].cpu())[2]
        loss_ref, grads_ref = O.siren_loss_and_grads(ref, grid.cpu(), img.cpu(), 50.0, 30.0)
        assert abs(l0 - loss_ref.item()) <= 2e-4 * loss_ref.item()
        assert _rel(model.layers[1].linear.weight.grad, grads_ref[2]) <= 1e-2
        for _ in range(3):
            th.train_epoch(model, optim, grid, img, lr_scheduler=sched)
    qm = q.convert()
    for layer in qm.layers:
        lin = layer.linear
        assert lin.weight_codes.dtype == torch.int8
        assert torch.equal(lin.weight.data, lin.weight_codes.float() * lin.weight_scales[:, None])


def test_deepcopy_and_rebinding_weight_data():
    """compress.py:174 deep-copies the model; masks/k-means rebind weight.data to new storage: pointers are
    read per call, never cached."""
    _, get_grid, _, Siren, _ = _pkg()
    torch.manual_seed(0)
    model = Siren(depth=3, hidden_size=128, first_omega_0=50, hidden_omega_0=30).cuda()
    grid = get_grid(16, 16, "cuda")
    with torch.no_grad():
        a = model(grid)
        m2 = copy.deepcopy(model)
        b = m2(grid)
        assert torch.equal(a, b)
        for p in m2.parameters():
            p.data = (p.data * 0.5).clone()  # new storage
        c = m2(grid)
        ref = O.siren_forward([p.detach().cpu() for p in m2.parameters()], grid.cpu(), 50.0, 30.0)
    assert (c.cpu() - ref).abs().max().item() <= 5e-4
    assert not torch.equal(a, c)


@pytest.mark.parametrize("hidden,depth,H,W", [(256, 5, 40, 56), (128, 4, 33, 47), (512, 4, 20, 36)])
def test_fused_kernel_variants_match_the_kernels_they_replace(monkeypatch, hidden, depth, H, W):
    """The fused kernels of the tensor-core path against the stand-alone kernels they replace (selected by
    environment switches read at handle creation), same weights, same image:
      tail kernel (last hidden GEMM + output layer + loss + dZ)  -> every gradient and the loss bit-identical
      layer 0 generated inside the first hidden GEMM              -> bit-identical (same arithmetic per element)
      layer-0 gradient reduced inside the first dX GEMM           -> dW0 / db0 differ by summation order only
    (with one sweep direction for every launch: SIRENB200_ALT_SWEEP=0), and
      launches sweeping the tiles in alternating directions       -> every reduction differs by summation order only
    """
    _, get_grid, synth_image, Siren, _ = _pkg()
    grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")

    def run(env):
        for k in ("SIRENB200_TAIL", "SIRENB200_GEN_FIRST", "SIRENB200_FUSE_L0", "SIRENB200_PDL"):
            monkeypatch.delenv(k, raising=False)
        monkeypatch.setenv("SIRENB200_ALT_SWEEP", "0")
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        torch.manual_seed(0)
        model = Siren(depth=depth, hidden_size=hidden, first_omega_0=50, hidden_omega_0=30,
                      precision="f16tc").cuda()
        eng = model.engine_for(grid)
        grads = [torch.zeros_like(p) for p in model.hot_parameters()]
        stats = torch.zeros(4, device="cuda")
        for _ in range(2):
            eng.forward_backward(model.kernel_parameters(), img, grads, stats)
        torch.cuda.synchronize()
        return [g.clone() for g in grads], stats.clone()

    base_g, base_s = run({})
    for env, exact_from in (({"SIRENB200_TAIL": "0"}, 0), ({"SIRENB200_GEN_FIRST": "0"}, 0),
                            ({"SIRENB200_PDL": "0"}, 0), ({"SIRENB200_FUSE_L0": "0"}, 2)):
        g, s = run(env)
        assert torch.equal(s[:2], base_s[:2]), env
        for i, (a, b) in enumerate(zip(g, base_g)):
            if i >= exact_from:
                assert torch.equal(a, b), (env, i)
            else:
                assert _rel(a, b) <= 1e-5, (env, i, _rel(a, b))
    g, s = run({"SIRENB200_ALT_SWEEP": "7"})
    assert abs(s[1].item() - base_s[1].item()) <= 1e-6 * abs(base_s[1].item())
    for i, (a, b) in enumerate(zip(g, base_g)):
        assert _rel(a, b) <= 1e-5, ("alternating sweep", i, _rel(a, b))


@pytest.mark.parametrize("n,seed", [(331_000, 0), (4097, 1), (31, 2), (65_536, 3)])
def test_device_prune_threshold_search_matches_the_reference_loop(n, seed):
    """sirenb200_prune_threshold_search (one warp, on the sorted magnitudes) follows the reference's multiplicative
    search (prune.py:74-95) step for step: same final threshold bit for bit (IEEE doubles), same removed count —
    including ties, zeros (already pruned weights), a NaN, and searches that end through the ten-stalls rule."""
    from implicit_image_compression_b200 import _lib
    from implicit_image_compression_b200.pipeline.masking.funcs.prune import _host_threshold_search
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(n, generator=g) * 0.05
    w[::5] = 0.0
    if n > 100:
        w[3::13] = w[7]       # ties
        w[11] = float("nan")
    mags = torch.sort(w.abs())[0]
    nonzero_total = int((w != 0).sum())  # NaN != 0 counts, as it does in the reference's mask statistics
    lib = _lib.load()
    for threshold, increment, frac, tol in [(1e-3, 0.2, 0.3, 0.02), (0.5, 0.2, 0.1, 0.0), (1e-6, 0.5, 0.6, 0.05),
                                            (0.02, 0.01, 0.25, 1e-4)]:
        tokill = max(1, math.ceil(frac * nonzero_total))
        want_t, want_removed = _host_threshold_search(mags.numpy(), nonzero_total, tokill, tol, threshold, increment)
        state = torch.tensor([threshold, increment, 0.0], dtype=torch.float64, device="cuda")
        dm = mags.cuda()
        _lib.check(lib.sirenb200_prune_threshold_search(dm.data_ptr(), dm.numel(), nonzero_total, tokill, tol,
                                                        state.data_ptr(), state.data_ptr() + 16,
                                                        torch.cuda.current_stream().cuda_stream))
        got_t, _, got_removed = state.tolist()
        assert got_t == want_t, (n, threshold, increment, frac, tol, got_t, want_t)
        assert int(got_removed) == want_removed
