"""GPU tests (-m gpu) of quant=qat with activation observers: the per-tensor fake-quant operator
against torch's own fused operator, and Quantize(QAT) end to end against vectors recorded from the unmodified
reference (tools/make_golden.py qat_case -> tests/golden/qat.npz)."""
import numpy as np
import pytest
import torch

import siren_oracle as O

pytestmark = pytest.mark.gpu


def test_act_fake_quant_operator_bit_exact_vs_torch():
    from implicit_image_compression_b200 import engine
    rng = np.random.RandomState(1)
    torch.manual_seed(1)
    for trial in range(25):
        n = int(rng.randint(1, 300000))
        x = torch.randn(n) * float(10 ** rng.uniform(-5, 1.5)) + float(rng.uniform(-1, 1))
        rmin, rmax = torch.tensor(float("inf")), torch.tensor(float("-inf"))
        scale, zp = torch.tensor([1.0]), torch.tensor([0], dtype=torch.int32)
        state = torch.tensor([float("inf"), float("-inf"), 1.0, 0.0], device="cuda")
        ostate = [float("inf"), float("-inf"), 1.0, 0.0]
        for it in range(3):
            xx = x * (1 + 0.25 * it) + 0.01 * it
            want = torch.fused_moving_avg_obs_fake_quant(xx, torch.tensor([1]), torch.tensor([1]), rmin, rmax, scale,
                                                         zp, 0.01, 0, 127, 0, False, False)
            got, mask = engine.fakequant_per_tensor(xx.cuda(), state, training=True, want_mask=True)
            _, omask, ostate = O.fused_obs_fake_quant(xx, ostate)
            assert torch.equal(got.cpu(), want), f"trial {trial} iteration {it}"
            assert torch.equal(mask.cpu().bool(), omask)
            st = state.tolist()
            assert st[0] == float(rmin) and st[1] == float(rmax) and st[2] == float(scale) and st[3] == float(zp)
        # frozen observers: state unchanged, same quantisation grid
        before = state.clone()
        got = engine.fakequant_per_tensor((x * 3).cuda(), state, training=False)
        assert torch.equal(state, before)
        want, _, _ = O.fused_obs_fake_quant(x * 3, ostate, training=False)
        assert torch.equal(got.cpu(), want)


def test_quantize_qat_end_to_end_vs_reference_golden(golden):
    from implicit_image_compression_b200.models import Siren
    from implicit_image_compression_b200.pipeline.quant.context import Quantize
    from implicit_image_compression_b200.utils import train_helper as th
    g = golden("qat.npz")
    torch.manual_seed(0)
    model = Siren(depth=4, hidden_size=32, first_omega_0=50, hidden_omega_0=30, precision="fp32")
    for i, p in enumerate(model.parameters()):
        assert torch.equal(p.detach(), torch.from_numpy(g[f"param{i}"]))
    model = model.cuda()
    grid, img = torch.from_numpy(g["grid"]).cuda(), torch.from_numpy(g["img"]).cuda()
    optim, sched = th.get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4}, quantize_mode=True)
    model.train()
    losses, evals = [], []
    with Quantize(model, optim, {"name": "QAT", "qconfig": "fbgemm", "num_steps": 6}) as q:
        for i in range(6):
            losses.append(th.train_epoch(model, optim, grid, img, lr_scheduler=sched))
            if i == 2:
                evals.append(th.eval_epoch(model, grid, img)[1])
                model.train()
        state = model._act_quant["state"].cpu().numpy()
        params_after = [p.detach().cpu().numpy() for p in model.parameters()]
    # 7-bit activations on a SIREN are coarse (one code of layer 0 moves the sine argument by ~0.7 rad), and which
    # code a borderline pre-activation rounds to depends on the fp32 summation order of the GEMM that produced it:
    # the trajectories agree to a fraction of a percent, not to fp32 accuracy
    np.testing.assert_allclose(losses, g["losses"], rtol=1.5e-2)
    np.testing.assert_allclose(evals, g["eval_losses"], rtol=1.5e-2)
    # layer 0's observer sees a deterministic input and slowly moving weights: tight; deeper layers inherit the
    # code flips of the layers before them
    for i in range(4):
        tol = 2e-2 if i == 0 else 0.15
        np.testing.assert_allclose(state[i, 0], g[f"act_min{i}"], rtol=tol, atol=1e-3)
        np.testing.assert_allclose(state[i, 1], g[f"act_max{i}"], rtol=tol, atol=1e-3)
        np.testing.assert_allclose(state[i, 2], g[f"act_scale{i}"], rtol=tol)
    for i, p in enumerate(params_after):
        np.testing.assert_allclose(p, g[f"param_after{i}"], rtol=0, atol=2 * 6 * 3e-4)  # <= 6 Adam steps of lr 3e-4
    observers = {n: (ob.min_val.clone(), ob.max_val.clone()) for n, ob in q._observers.items()}
    masters = {n: m.weight.detach().clone() for n, m in q._targets}
    qm = q.convert()
    assert not qm.training
    for i, layer in enumerate(qm.layers):
        lin, name = layer.linear, f"layers.{i}.linear"
        # convert(): the int8 tensors are those of torch's per-channel observer formula on the final master weights
        # (an Adam step moves a hidden weight by ~3 int8 codes here, so the codes of two trajectories that differ
        # in a few activation codes are not comparable element by element; shapes, ranges and scales are)
        lo, hi = observers[name]
        codes, scales, deq = O.fake_quant_per_channel_weight(masters[name].cpu(), lo.cpu(), hi.cpu())
        assert torch.equal(lin.weight_codes.cpu(), codes)
        assert torch.equal(lin.weight_scales.cpu(), scales)
        assert torch.equal(lin.weight.data.cpu(), deq)
        want = g[f"int8_w{i}"]
        assert tuple(lin.weight_codes.shape) == want.shape and lin.weight_codes.dtype == torch.int8
        np.testing.assert_allclose(lin.weight_scales.cpu().numpy(), g[f"int8_w_scale{i}"], rtol=0.1)
        np.testing.assert_allclose(float(lin.act_scale), float(g[f"int8_out_scale{i}"]), rtol=0.15)
    # the converted model evaluates with frozen observers: two forwards give the same prediction
    with torch.no_grad():
        a = qm(grid).clone()
        b = qm(grid)
    assert torch.equal(a, b)
