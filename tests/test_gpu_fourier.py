"""GPU tests (-m gpu) of the second registry model, FourierNet (reference models/fourier.py), against
vectors recorded from the unmodified reference (tests/golden/fourier.npz) and the oracle."""
import numpy as np
import pytest
import torch

import siren_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _model(g):
    from implicit_image_compression_b200.models import registry
    torch.manual_seed(0)
    model = registry["fourier"](name="fourier", depth=int(g["depth"]), hidden_size=int(g["hidden"]),
                                map_size=int(g["map_size"]), map_scale=float(g["map_scale"]))
    for n, p in model.named_parameters():
        assert torch.equal(p.detach(), torch.from_numpy(g["param/" + n])), f"init RNG stream differs: {n}"
    return model.cuda()


def test_fouriernet_forward_loss_grads_vs_reference_golden(golden):
    g = golden("fourier.npz")
    model = _model(g)
    grid, img = torch.from_numpy(g["grid"]).cuda(), torch.from_numpy(g["img"]).cuda()
    with torch.no_grad():
        pred = model(grid)
    # (the encoding's sine arguments reach ~300 rad, where one fp32 ulp is 3e-5)
    assert (pred.cpu() - torch.from_numpy(g["pred"])).abs().max().item() <= 2e-4
    params = model.hot_parameters()
    grads = [torch.empty_like(p) for p in params]
    stats = model.engine_for(grid).forward_backward(model.kernel_parameters(), img, grads).tolist()
    assert abs(stats[1] - float(g["loss"])) <= 1e-4 * float(g["loss"])
    names = [n for n, p in model.named_parameters() if n != "encoding.B"]
    for n, gr in zip(names, grads):
        assert _rel(gr, torch.from_numpy(g["grad/" + n])) <= 2e-3, n
    # autograd bridge (model(grid) under grad mode + loss.backward()) gives the same gradients
    loss = torch.nn.functional.mse_loss(model(grid), img)
    loss.backward()
    for p, gr in zip(params, grads):
        assert _rel(p.grad, gr) <= 1e-5


def test_fouriernet_fit_matches_reference_trajectory(golden):
    from implicit_image_compression_b200.utils import train_helper as th
    g = golden("fourier.npz")
    model = _model(g)
    grid, img = torch.from_numpy(g["grid"]).cuda(), torch.from_numpy(g["img"]).cuda()
    optim, sched = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
    losses = [th.train_epoch(model, optim, grid, img, lr_scheduler=sched) for _ in range(10)]
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-4)
    _, l_, psnr, _ = th.eval_epoch(model, grid, img)
    assert abs(l_ - float(g["eval_loss"])) <= 5e-4 * float(g["eval_loss"])
    assert abs(psnr - float(g["eval_psnr"])) <= 5e-3


def test_fouriernet_c2_size_vs_oracle_band():
    """Full config-2 image: prediction of a 16-row band against the oracle, and a Fitter run (graph-free fp32 path)."""
    from implicit_image_compression_b200.data import get_grid, synth_image
    from implicit_image_compression_b200.fit import Fitter
    from implicit_image_compression_b200.models import FourierNet
    from implicit_image_compression_b200.utils import train_helper as th
    H, W = 128, 192
    torch.manual_seed(0)
    model = FourierNet(depth=4, hidden_size=64, map_size=64, map_scale=10.0)
    B = model.encoding.B.detach().clone()
    ref = [p.detach().clone() for p in model.hot_parameters()]
    model = model.cuda()
    grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")
    with torch.no_grad():
        pred = model(grid)
    band = slice(40, 56)
    want = O.fourier_forward(B, ref, grid[band].cpu())
    assert (pred[band].cpu() - want).abs().max().item() <= 2e-4
    optim, sched = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
    losses = Fitter(model, optim, grid, img, sched).steps(30).tolist()
    assert losses[-1] < losses[0]
