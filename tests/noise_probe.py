#!/usr/bin/env python
"""Forward / gradient noise of the f16tc data path AT TRAINED WEIGHTS (checker script for the GPU box).

A fit is run with the product path to several checkpoints; at each one the fp32 master weights are evaluated with
  * the exact fp32 torch forward (reference math),
  * the library's f16tc and fp32 forwards,
  * torch emulations of the f16tc forward with selected roundings,
and the gradient of the library / the emulations is compared with autograd's (relative L2 error and cosine per
tensor).  This separates "the rounded forward is a noisier function" from "the rounded gradient misleads Adam".
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import psnr_study as P  # noqa: E402


def psnr(pred, img):
    return 10 * math.log10(1 / torch.mean((pred - img) ** 2).item())


@torch.no_grad()
def emu_forward(params, x, act, w16):
    depth = len(params) // 2
    rw = P.q16 if w16 else (lambda t: t)
    t = P.OMEGA0 * torch.addmm(params[1], x, params[0].t())
    for l in range(depth - 1):
        a = torch.sin(t)
        if act == "sh":
            a = P.signed_half(a, torch.cos(t) < 0)
        elif act == "h":
            a = P.q16(a)
        elif act == "h0":      # only layer 0's output kept exact, the rest fp16
            a = a if l == 0 else P.q16(a)
        w, b = params[2 * (l + 1)], params[2 * (l + 1) + 1]
        if l == depth - 2:
            return torch.addmm(b, a, rw(w).t()) / 2 + 0.5
        t = P.OMEGA * torch.addmm(b, a, rw(w).t())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--image", type=int, default=0)
    ap.add_argument("--checkpoints", default="250,1000,2000,4000")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "noise_probe.jsonl"))
    args = ap.parse_args()
    from implicit_image_compression_b200.data import get_grid, synth_image
    from implicit_image_compression_b200.fit import Fitter
    from implicit_image_compression_b200.models import Siren
    from implicit_image_compression_b200.utils.train_helper import get_optimizer_lr_scheduler
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    H, W = 512, 768
    grid, img = get_grid(H, W, dev), synth_image(H, W, args.image, device=dev)
    x = ((grid.view(-1, 2) - 0.5) * 2).contiguous()
    tgt = img.view(-1, 3)
    torch.manual_seed(0)
    model = Siren(depth=6, hidden_size=256, first_omega_0=P.OMEGA0, hidden_omega_0=P.OMEGA, precision="f16tc").to(dev)
    m32 = Siren(depth=6, hidden_size=256, first_omega_0=P.OMEGA0, hidden_omega_0=P.OMEGA, precision="fp32").to(dev)
    optim, sched = get_optimizer_lr_scheduler(model, {"name": "adam", "lr": P.LR})
    fitter = Fitter(model, optim, grid, img, sched)
    done = 0
    with open(args.out, "a") as f:
        for ck in [int(c) for c in args.checkpoints.split(",")]:
            fitter.steps(ck - done)
            done = ck
            torch.cuda.synchronize()
            params = [p.detach().clone() for p in model.hot_parameters()]
            row = {"image": args.image, "step": ck}
            with torch.no_grad():
                exact = P.fwd32(params, x)
                row["psnr_exact_fp32"] = psnr(exact, tgt)
                model.eval()
                row["psnr_lib_f16tc"] = psnr(model(grid).view(-1, 3), tgt)
                model.train()
                for p, q in zip(m32.hot_parameters(), params):
                    p.data.copy_(q)
                row["psnr_lib_fp32"] = psnr(m32(grid).view(-1, 3), tgt)
                for name, (act, w16) in {"emu_sh_w16": ("sh", True), "emu_h_w16": ("h", True),
                                         "emu_h_w32": ("h", False), "emu_f_w16": ("f", True),
                                         "emu_h0_w16": ("h0", True)}.items():
                    pr = emu_forward(params, x, act, w16)
                    row["psnr_" + name] = psnr(pr, tgt)
                    row["fnoise_rms_" + name] = torch.sqrt(torch.mean((pr - exact) ** 2)).item()
                row["fnoise_rms_lib_f16tc"] = torch.sqrt(torch.mean((model(grid).view(-1, 3) - exact) ** 2)).item()
                row["fit_rmse_exact"] = torch.sqrt(torch.mean((exact - tgt) ** 2)).item()
                row["weight_rms"] = [p.pow(2).mean().sqrt().item() for p in params[0::2]]
            # gradients
            ref = [p.clone().requires_grad_(True) for p in params]
            torch.nn.functional.mse_loss(P.fwd32(ref, x), tgt).backward()
            gref = [p.grad for p in ref]

            def cmp(grads):
                rel = [((g.reshape(r.shape) - r).norm() / (r.norm() + 1e-30)).item() for g, r in zip(grads, gref)]
                cos = [torch.nn.functional.cosine_similarity(g.reshape(-1), r.reshape(-1), dim=0).item()
                       for g, r in zip(grads, gref)]
                return {"rel_l2": [round(v, 5) for v in rel], "cos": [round(v, 6) for v in cos]}

            grads = [torch.empty_like(p) for p in params]
            stats = fitter.engine.forward_backward(params, img, grads)
            row["grad_lib_f16tc"] = cmp(grads)
            with torch.no_grad():
                for name in ("emu_cur", "emu_h", "emu_fwd32", "emu_bwd32"):
                    st = {"G": 2.0 ** math.floor(math.log2(0.125 / row["fit_rmse_exact"]))}
                    _, g = P.emu_loss_and_grads(params, x, tgt, P.EMU[name], st)
                    row["grad_" + name] = cmp(g)
            print(json.dumps(row), flush=True)
            f.write(json.dumps(row) + "\n")
            # restore the fitter's own gradient state is not needed: forward_backward overwrites everything


if __name__ == "__main__":
    main()
