"""Quick GPU parity check against the oracle (developer tool; the formal version is tests/ -m gpu)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import siren_oracle as O  # noqa: E402
from implicit_image_compression_b200 import engine as E  # noqa: E402
from implicit_image_compression_b200.data import get_grid, synth_image  # noqa: E402
from implicit_image_compression_b200.models import Siren  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def run(depth, hidden, H, W, precision):
    torch.manual_seed(0)
    model = Siren(depth=depth, hidden_size=hidden, first_omega_0=50, hidden_omega_0=30,
                  precision=precision)
    ref_params = O.siren_init(0, depth, hidden, 50.0, 30.0)
    for p, r in zip(model.parameters(), ref_params):
        assert torch.equal(p.detach(), r), "init mismatch"
    grid_cpu = O.get_grid(H, W)
    img_cpu = O.synth_image(H, W, 0)
    assert torch.equal(grid_cpu, get_grid(H, W))
    assert torch.equal(img_cpu, synth_image(H, W, 0))
    pred_ref = O.siren_forward(ref_params, grid_cpu, 50.0, 30.0)
    loss_ref, grads_ref = O.siren_loss_and_grads(ref_params, grid_cpu, img_cpu, 50.0, 30.0)

    model = model.cuda()
    grid, img = grid_cpu.cuda(), img_cpu.cuda()
    with torch.no_grad():
        pred = model(grid)
    torch.cuda.synchronize()
    print(f"[{precision} D{depth} W{hidden} {H}x{W}] pred max|d| = {(pred.cpu() - pred_ref).abs().max():.3e}")
    params = model.hot_parameters()
    grads = [torch.zeros_like(p) for p in params]
    eng = model.engine_for(grid)
    stats = eng.forward_backward([p.data for p in params], img, grads)
    torch.cuda.synchronize()
    s = stats.tolist()
    print(f"   loss {s[1]:.8f} ref {loss_ref.item():.8f}  nonfinite {s[2]}")
    for i, (g, r) in enumerate(zip(grads, grads_ref)):
        print(f"   grad[{i}] rel err {rel(g, r):.3e}  |ref| {r.norm():.3e}")
    # second step exercises the seed-scale update
    stats = eng.forward_backward([p.data for p in params], img, grads)
    torch.cuda.synchronize()
    print(f"   2nd call: loss {stats[1].item():.8f}  grad[2] rel {rel(grads[2], grads_ref[2]):.3e}")
    # autograd path
    pred2 = model(grid)
    loss2 = torch.nn.functional.mse_loss(pred2, img)
    loss2.backward()
    torch.cuda.synchronize()
    for i, (p, r) in enumerate(zip(params, grads_ref)):
        if i in (0, 2, len(grads_ref) - 2):
            print(f"   autograd grad[{i}] rel err {rel(p.grad, r):.3e}")


def misc():
    torch.manual_seed(1)
    # Adam
    p = torch.randn(1000)
    g = torch.randn(1000) * 1e-3
    m = torch.zeros(1000)
    v = torch.zeros(1000)
    pr, mr, vr = p.clone(), m.clone(), v.clone()
    pc, gc, mc, vc = p.cuda(), g.cuda(), m.cuda(), v.cuda()
    for step in range(1, 4):
        pr, mr, vr = O.adam_step(pr, g, mr, vr, step, 3e-4)
        E.adam_step([pc], [gc], [mc], [vc], None, 3e-4, 0.9, 0.999, 1e-8, step)
    torch.cuda.synchronize()
    print(f"adam: p max|d| {(pc.cpu() - pr).abs().max():.3e}  m {(mc.cpu() - mr).abs().max():.3e}")
    # mask
    w = torch.randn(64, 64)
    mk = (torch.rand(64, 64) < 0.5).float()
    wc = w.cuda()
    E.apply_mask_(wc, mk.cuda())
    print("mask bit-exact:", torch.equal(wc.cpu(), O.apply_mask(w, mk)))
    # kmeans
    w = (torch.rand(128, 128) - 0.5) * 0.02
    w[torch.rand(128, 128) < 0.3] = 0
    for bits in (4, 8):
        cr, lr_, wr = O.kmeans_quantize(w, bits)
        nz = w[w != 0]
        init = torch.linspace(nz.min().item(), nz.max().item(), 2 ** bits - 1)
        t0 = time.time()
        cg, lg, wg = E.kmeans_quantize(w.cuda(), bits, init_centers=init.cuda())
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"kmeans bits={bits}: ncent {cg.numel()} vs {cr.numel()}  centroids equal "
              f"{cg.numel() == cr.numel() and torch.equal(cg.cpu(), cr)}  labels equal "
              f"{torch.equal(lg.cpu(), lr_)}  weights equal {torch.equal(wg.cpu(), wr)}  ({dt*1e3:.1f} ms)")
    # fake quant
    w = torch.randn(32, 100) * 0.01
    qr, sr, dr = O.fake_quant_per_channel_weight(w)
    qg, sg, dg = E.fakequant_per_channel(w.cuda())
    print("fakequant codes equal:", torch.equal(qg.cpu(), qr), " scales equal:", torch.equal(sg.cpu(), sr),
          " deq equal:", torch.equal(dg.cpu(), dr))
    # eval metrics
    a, b = torch.rand(50, 60, 3), torch.rand(50, 60, 3)
    mse, psnr, psnr8 = O.eval_metrics(a, b)
    mg = E.eval_metrics(a.cuda(), b.cuda()).tolist()
    print(f"eval: mse {mg[0]:.8f} vs {mse:.8f}; mse8 -> psnr8 ref {psnr8:.4f}")


if __name__ == "__main__":
    misc()
    run(3, 64, 40, 48, "fp32")
    run(4, 128, 40, 48, "fp32")
    run(4, 128, 40, 48, "f16tc")
    run(6, 256, 96, 128, "f16tc")
    run(2, 128, 33, 47, "f16tc")
    print("gpu_check done")
