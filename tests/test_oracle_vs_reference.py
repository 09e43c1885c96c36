"""Live check of the oracle against the reference (only where /root/reference exists, i.e. the build
container); the GPU box relies on the committed golden vectors instead."""
import pytest
import torch

import ref_import
import siren_oracle as O

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def test_forward_backward_matches_reference_autograd():
    ns = ref_import.load()
    torch.manual_seed(3)
    m = ns.model_registry["siren"](name="siren", depth=5, hidden_size=24, first_omega_0=50, hidden_omega_0=30,
                                   outermost_linear=True, simulate_quantization=False,
                                   small_dense_density=1.0)
    params = O.siren_init(3, 5, 24, 50.0, 30.0)
    assert all(torch.equal(a.detach(), b) for a, b in zip(m.parameters(), params))
    grid = ns.get_grid(9, 13)
    assert torch.equal(grid, O.get_grid(9, 13))
    img = O.synth_image(9, 13, 4)
    loss = torch.nn.functional.mse_loss(m(grid), img)
    loss.backward()
    l2, grads = O.siren_loss_and_grads(params, grid, img, 50.0, 30.0)
    assert abs(loss.item() - l2.item()) < 1e-8
    for p, g in zip(m.parameters(), grads):
        assert (p.grad - g).norm() <= 2e-6 * p.grad.norm() + 1e-12


def test_sine_last_layer_variant():
    ns = ref_import.load()
    torch.manual_seed(1)
    m = ns.model_registry["siren"](name="siren", depth=3, hidden_size=8, first_omega_0=50, hidden_omega_0=30,
                                   outermost_linear=False, simulate_quantization=False,
                                   small_dense_density=1.0)
    params = [p.detach().clone() for p in m.parameters()]
    grid, img = ns.get_grid(6, 5), O.synth_image(6, 5, 0)
    loss = torch.nn.functional.mse_loss(m(grid), img)
    loss.backward()
    l2, grads = O.siren_loss_and_grads(params, grid, img, 50.0, 30.0, outermost_linear=False)
    assert abs(loss.item() - l2.item()) < 1e-8
    for p, g in zip(m.parameters(), grads):
        assert (p.grad - g).norm() <= 2e-6 * p.grad.norm() + 1e-12


def test_eval_metrics_match_reference():
    ns = ref_import.load()
    torch.manual_seed(0)
    m = ns.model_registry["siren"](name="siren", depth=3, hidden_size=8, first_omega_0=50, hidden_omega_0=30,
                                   outermost_linear=True, simulate_quantization=False,
                                   small_dense_density=1.0)
    grid, img = ns.get_grid(7, 7), O.synth_image(7, 7, 0)
    pred, loss, psnr, psnr8 = ns.eval_epoch(m, grid, img)
    a, b, c = O.eval_metrics(pred, img)
    assert abs(a - loss) < 1e-9 and abs(b - psnr) < 1e-5 and abs(c - psnr8) < 1e-5
