"""GPU tests (-m gpu): GradScaler semantics with a REAL overflow, sparsity preserved through
the quantisation fine-tune, optimizer checkpoint round trip.

The yardstick for the scaler is torch's own GradScaler driving the reference's math (nn.Linear + sin + mse_loss +
torch.optim.Adam, i.e. utils/train_helper.py:132-185 with `scaler`) on the same GPU."""
import copy

import pytest
import torch
from torch.nn import functional as F

pytestmark = pytest.mark.gpu

OMEGA0, OMEGA = 50.0, 30.0


def _pkg():
    from implicit_image_compression_b200.data import get_grid, synth_image
    from implicit_image_compression_b200.models import Siren
    from implicit_image_compression_b200.utils import train_helper
    return get_grid, synth_image, Siren, train_helper


def _torch_forward(params, grid):
    depth = len(params) // 2
    a = ((grid.view(-1, 2) - 0.5) * 2)
    for l in range(depth):
        z = F.linear(a, params[2 * l], params[2 * l + 1])
        if l == depth - 1:
            return (z / 2 + 0.5).view(grid.shape[0], grid.shape[1], -1)
        a = torch.sin((OMEGA0 if l == 0 else OMEGA) * z)


@pytest.mark.parametrize("precision,hidden", [("fp32", 32), ("f16tc", 128)])
def test_gradscaler_real_overflow_matches_torch(precision, hidden):
    get_grid, synth_image, Siren, th = _pkg()
    H, W, depth = 40, 56, 4
    torch.manual_seed(0)
    model = Siren(depth=depth, hidden_size=hidden, first_omega_0=OMEGA0, hidden_omega_0=OMEGA,
                  precision=precision).cuda()
    grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")
    ref = [p.detach().clone().requires_grad_(True) for p in model.hot_parameters()]
    ref_opt = torch.optim.Adam(ref, lr=3e-4)
    ref_scaler = torch.amp.GradScaler("cuda", init_scale=2.0 ** 16, growth_interval=3)
    optim, sched = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
    scaler = torch.amp.GradScaler("cuda", init_scale=2.0 ** 16, growth_interval=3)
    tol = 1e-5 if precision == "fp32" else 5e-2

    def ref_step():
        ref_opt.zero_grad()
        loss = F.mse_loss(_torch_forward(ref, grid), img)
        ref_scaler.scale(loss).backward()
        ref_scaler.step(ref_opt)
        ref_scaler.update()
        return loss.item()

    for step in range(8):
        if step == 4:  # a REAL overflow: the output bias is so large that pred, the error and every gradient are inf
            with torch.no_grad():
                saved = model.layers[-1].linear.bias.detach().clone()
                model.layers[-1].linear.bias.fill_(3.0e38)
                ref[-1].fill_(3.0e38)
        before = [p.detach().clone() for p in model.hot_parameters()]
        ref_before = [p.detach().clone() for p in ref]
        loss = th.train_epoch(model, optim, grid, img, lr_scheduler=sched, scaler=scaler)
        ref_loss = ref_step()
        assert scaler.get_scale() == ref_scaler.get_scale(), f"step {step}: scale differs"
        assert int(scaler._growth_tracker.item()) == int(ref_scaler._growth_tracker.item())
        if step == 4:
            assert not torch.isfinite(torch.tensor(loss)) and not torch.isfinite(torch.tensor(ref_loss))
            # both skipped the optimizer step
            assert all(torch.equal(a, b.detach()) for a, b in zip(before, model.hot_parameters()))
            assert all(torch.equal(a, b.detach()) for a, b in zip(ref_before, ref))
            with torch.no_grad():
                model.layers[-1].linear.bias.copy_(saved)
                ref[-1].copy_(saved)
        else:
            assert abs(loss - ref_loss) <= max(1e-4, tol) * abs(ref_loss) + 1e-7
            for p, r in zip(model.hot_parameters(), ref):
                # unscaled gradients are what both leave in .grad
                err = (p.grad - r.grad).norm() / (r.grad.norm() + 1e-30)
                assert err.item() <= tol, f"step {step}: grad mismatch {err.item():.3e}"
    # the optimizer step count did not advance on the skipped step: 7 real steps
    assert optim.param_groups[0]["_fused_step"] == 7
    assert int(ref_opt.state[ref[0]]["step"]) == 7


def test_gradscaler_with_mask_skips_only_the_optimizer():
    get_grid, synth_image, Siren, th = _pkg()
    H, W = 32, 48
    torch.manual_seed(0)
    model = Siren(depth=4, hidden_size=32, first_omega_0=OMEGA0, hidden_omega_0=OMEGA, precision="fp32").cuda()
    grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")
    optim, sched = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
    cfg = dict(name="RigL", dense=False, density=0.5, sparse_init="random", dense_gradients=False,
               prune_mode="magnitude", growth_mode="absolute-gradient", redistribution_mode="none",
               decay_schedule="cosine", prune_rate=0.3, end_when=100, interval=10, start_when=0, final_density=0.5)
    mask = th.setup_mask(model, optim, cfg)
    scaler = torch.amp.GradScaler("cuda")
    th.train_epoch(model, optim, grid, img, lr_scheduler=sched, mask=mask, scaler=scaler)
    ms0 = mask.mask_step
    with torch.no_grad():
        model.layers[-1].linear.bias.fill_(3.0e38)
    before = [p.detach().clone() for p in model.hot_parameters()]
    th.train_epoch(model, optim, grid, img, lr_scheduler=sched, mask=mask, scaler=scaler)
    assert all(torch.equal(a, b.detach()) for a, b in zip(before, model.hot_parameters()))
    assert mask.mask_step == ms0 + 1            # core.py:702 runs whether or not the optimizer stepped
    assert scaler.get_scale() == 2.0 ** 15      # backoff
    for n, w in mask._masked_parameters():      # masks still applied
        assert torch.equal(w.detach() * mask.mask_dict[n], w.detach())


def test_quant_phase_keeps_pruned_weights_zero():
    """compress.main: Pruning then 8-bit k-means fine-tune; the quantised copy must keep the mask's zeros
    (the fine-tune used to step pruned weights away from zero)."""
    from implicit_image_compression_b200.compress import main
    from implicit_image_compression_b200.config import load_config
    cfg = load_config(["mlp.hidden_size=128", "mlp.depth=4", "img.height=64", "img.width=96", "masking=Pruning",
                       "masking.final_density=0.3", "masking.end_when=60", "masking.interval=10",
                       "masking.start_when=5", "train.num_steps=80", "train.multiplier=1", "train.log_steps=40",
                       "quant=kmeans", "quant.bits=5", "quant.num_steps=6", "quant.log_steps=3"])
    cfg.quant["skip_ll"] = ["layers.0.linear", "layers.3.linear"]
    out = main(cfg)
    assert abs(out["Density"] - out["Quant Density"]) <= 0.02, out
    assert out["Quant Density"] <= 0.5
    cfg2 = copy.deepcopy(cfg)
    cfg2.quant["replicate_reference_mask_bug"] = True
    out2 = main(cfg2)
    assert out2["quant_replicates_reference_mask_bug"] is True
    assert out2["Quant Density"] <= 0.5


def test_fused_adam_loads_stock_adam_checkpoint():
    get_grid, synth_image, Siren, th = _pkg()
    torch.manual_seed(0)
    model = Siren(depth=3, hidden_size=32, precision="fp32").cuda()
    grid, img = get_grid(16, 16, "cuda"), synth_image(16, 16, 0, device="cuda")
    optim, _ = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
    for _ in range(5):
        th.train_epoch(model, optim, grid, img)
    sd = copy.deepcopy(optim.state_dict())
    for g in sd["param_groups"]:
        g.pop("_fused_step", None)  # what a stock torch.optim.Adam checkpoint looks like
    optim2, _ = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
    optim2.load_state_dict(sd)
    assert optim2.param_groups[0]["_fused_step"] == 5
