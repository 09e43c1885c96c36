#!/usr/bin/env python
"""Long-horizon PSNR parity study (SURVEY.md §8d "time-to-PSNR") — a CHECKER script, run on
the GPU box:  python tests/psnr_study.py --images 0-3 --steps 2000 --arms ref32,ref32_b,lib_f16tc

Yardstick arm `ref32`: the reference's own math in plain PyTorch on the same GPU — nn.Linear-style fp32 GEMMs
(TF32 off), torch.sin, autograd, F.mse_loss, torch.optim.Adam(3e-4) + StepLR(2000, 0.5) — i.e. models/siren.py:
56-68,123-134 + utils/train_helper.py:132-185 restated functionally (it is also the "eager PyTorch on the same
box" rate).  `ref32_b` is the same computation with the loss accumulated over two row halves: identical
mathematics, different fp32 summation order -> the yardstick's own run-to-run spread (self-noise).

`lib_*` arms run the product path (libsirenb200 through Fitter).  `emu_*` arms re-state the f16tc data path in
torch with explicit operand rounding, to locate which rounding the fitted PSNR is sensitive to.

Every arm is evaluated the same way: the fp32 torch forward of the arm's CURRENT fp32 master weights (what a
decoder of the compressed image would compute), every `--eval-every` steps over the last `--tail` steps; the
figure of merit is the median PSNR over that trailing window, averaged over images.
"""
import argparse
import json
import math
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OMEGA0, OMEGA, LR = 50.0, 30.0, 3e-4


def init_params(depth, hidden, dev):
    from implicit_image_compression_b200.models import Siren
    torch.manual_seed(0)
    m = Siren(depth=depth, hidden_size=hidden, first_omega_0=OMEGA0, hidden_omega_0=OMEGA)
    return [p.detach().clone().to(dev) for p in m.hot_parameters()]


def fwd32(params, x):
    """siren.py:123-134 in fp32: x in [-1,1]^2 [N,2] -> pred [N,3]."""
    depth = len(params) // 2
    a = x
    for l in range(depth):
        z = torch.addmm(params[2 * l + 1], a, params[2 * l].t())
        if l == depth - 1:
            return z / 2 + 0.5
        a = torch.sin((OMEGA0 if l == 0 else OMEGA) * z)


@torch.no_grad()
def psnr32(params, x, img):
    mse = torch.mean((fwd32(params, x) - img) ** 2).item()
    return 10 * math.log10(1 / mse)


# ------------------------------------------------------------------------------------------------------
# emulation of the f16tc data path (tc_kernels.cuh) with selectable roundings
# ------------------------------------------------------------------------------------------------------
def q16(t):
    return t.half().float()


def signed_half(a, cos_neg):
    """fp16(a) with the mantissa LSB replaced by [cos < 0] (tc_kernels.cuh sine_signed_half2)."""
    h = a.half().view(torch.int16)
    h = (h & ~1) | cos_neg.to(torch.int16)
    return h.view(torch.half).float()


def parity_half(a, cos_neg):
    """nearest fp16 whose mantissa LSB equals [cos < 0]."""
    h = a.half()
    bits = h.view(torch.int16)
    ok = (bits & 1) == cos_neg.to(torch.int16)
    up = (bits + 1).view(torch.half).float()     # next magnitude up (same sign)
    dn = (bits - 1).view(torch.half).float()
    hf = h.float()
    pick_up = (up - a).abs() <= (dn - a).abs()
    alt = torch.where(pick_up, up, dn)
    return torch.where(ok, hf, alt)


def emu_loss_and_grads(params, x, img, opt, state):
    """forward + MSE + backward with the roundings of the tensor-core path.  opt keys:
    act: 'sh' (signed-half, 9 bit + sign), 'par' (parity-rounded), 'h' (fp16), 'f' (fp32)
    w16: round GEMM weights to fp16;  dz16: round dZ / seed to fp16 (with the power-of-two seed scale G)
    cos: 'stash' (+-sqrt(1 - a_stash^2)) or 'exact'."""
    depth = len(params) // 2
    nh = depth - 2
    N = x.shape[0]
    rw = q16 if opt["w16"] else (lambda t: t)
    rdz = q16 if opt["dz16"] else (lambda t: t)
    G = state.get("G", 1.0) if opt["dz16"] else 1.0
    stash, cosx = [], []
    t = OMEGA0 * torch.addmm(params[1], x, params[0].t())
    a = torch.sin(t)
    for l in range(0, nh + 1):
        cneg = torch.cos(t) < 0
        if opt["act"] == "sh":
            s = signed_half(a, cneg)
        elif opt["act"] == "par":
            s = parity_half(a, cneg)
        elif opt["act"] == "h":
            s = q16(a)
        else:
            s = a
        stash.append(s)
        if opt["cos"] == "exact":
            cosx.append(torch.cos(t))
        else:
            c = torch.sqrt(torch.clamp(1 - s * s, min=0))
            cosx.append(torch.where(cneg, -c, c))
        if l == nh:
            break
        t = OMEGA * torch.addmm(params[2 * (l + 1) + 1], s, rw(params[2 * (l + 1)]).t())
        a = torch.sin(t)
        del cneg
    wl = params[2 * (depth - 1)]
    y = torch.addmm(params[2 * (depth - 1) + 1], stash[nh], rw(wl).t())
    pred = y / 2 + 0.5
    d = pred - img
    sse = (d * d).sum()
    loss = sse / d.numel()
    g = rdz(d * G)                                    # seed (the 1/2 of y/2 and 2/(3N) are folded into `scale`)
    scale = (2.0 * 0.5 / d.numel()) / G
    grads = [None] * (2 * depth)
    grads[2 * (depth - 1)] = (g.t() @ stash[nh]) * scale
    grads[2 * (depth - 1) + 1] = g.sum(0) * scale
    om_last = OMEGA if nh >= 1 else OMEGA0
    dz = rdz((g @ rw(wl * om_last)) * cosx[nh])       # omega of the cos factor folded into the weights
    for l in range(nh, 0, -1):
        grads[2 * l] = (dz.t() @ stash[l - 1]) * scale
        grads[2 * l + 1] = dz.sum(0) * scale
        om_prev = OMEGA0 if l - 1 == 0 else OMEGA
        dz = rdz((dz @ rw(params[2 * l] * om_prev)) * cosx[l - 1])
    grads[0] = (dz.t() @ x) * scale
    grads[1] = dz.sum(0) * scale
    if opt["dz16"]:
        lv = loss.item()
        g2 = 2.0 ** math.floor(math.log2(0.125 / math.sqrt(lv))) if lv > 0 else 1.0
        state["G"] = min(max(g2, 1.0), 4096.0)
    return loss, grads


EMU = {
    "emu_none": dict(act="f", w16=False, dz16=False, cos="exact"),
    "emu_cur": dict(act="sh", w16=True, dz16=True, cos="stash"),
    "emu_par": dict(act="par", w16=True, dz16=True, cos="stash"),
    "emu_h": dict(act="h", w16=True, dz16=True, cos="stash"),
    "emu_h_w32": dict(act="h", w16=False, dz16=True, cos="stash"),
    "emu_f_w16": dict(act="f", w16=True, dz16=True, cos="stash"),
    "emu_fwd32": dict(act="f", w16=False, dz16=True, cos="exact"),
    "emu_bwd32": dict(act="h", w16=True, dz16=False, cos="exact"),
}


# ------------------------------------------------------------------------------------------------------
def run_arm(arm, idx, args, dev):
    from implicit_image_compression_b200.data import get_grid, synth_image
    H, W = args.height, args.width
    grid = get_grid(H, W, dev)
    img = synth_image(H, W, idx, device=dev)
    x = ((grid.view(-1, 2) - 0.5) * 2).contiguous()
    tgt = img.view(-1, 3)
    evals = {}

    def want_eval(s):
        return s in (100, 200, 500, 1000, 1500) or (s > args.steps - args.tail and (args.steps - s) % args.eval_every == 0)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_eval = 0.0
    losses = []
    if arm.startswith("lib_"):
        from implicit_image_compression_b200.fit import Fitter
        from implicit_image_compression_b200.models import Siren
        from implicit_image_compression_b200.utils.train_helper import get_optimizer_lr_scheduler
        torch.manual_seed(0)
        model = Siren(depth=args.depth, hidden_size=args.hidden, first_omega_0=OMEGA0, hidden_omega_0=OMEGA,
                      precision=arm[4:]).to(dev)
        optim, sched = get_optimizer_lr_scheduler(model, {"name": "adam", "lr": LR})
        fitter = Fitter(model, optim, grid, img, sched)
        done = 0
        stops = sorted(s for s in range(1, args.steps + 1) if want_eval(s))
        for s in stops:
            ls = fitter.steps(s - done)
            done = s
            torch.cuda.synchronize()
            te = time.perf_counter()
            evals[s] = psnr32([p.detach() for p in model.hot_parameters()], x, tgt)
            t_eval += time.perf_counter() - te
        losses = [float(ls[-1])]
        del fitter, model
    else:
        params = [p.requires_grad_(True) for p in init_params(args.depth, args.hidden, dev)]
        optim = torch.optim.Adam(params, lr=LR)
        sched = torch.optim.lr_scheduler.StepLR(optim, 2000, gamma=0.5)
        state = {}
        half = (H // 2) * W
        for s in range(1, args.steps + 1):
            if arm == "ref32":
                optim.zero_grad(set_to_none=True)
                loss = torch.nn.functional.mse_loss(fwd32(params, x), tgt)
                loss.backward()
            elif arm == "ref32_b":
                optim.zero_grad(set_to_none=True)
                n_el = tgt.numel()
                l1 = ((fwd32(params, x[:half]) - tgt[:half]) ** 2).sum() / n_el
                l1.backward()
                l2 = ((fwd32(params, x[half:]) - tgt[half:]) ** 2).sum() / n_el
                l2.backward()
                loss = l1.detach() + l2.detach()
            else:
                with torch.no_grad():
                    loss, grads = emu_loss_and_grads(params, x, tgt, EMU[arm], state)
                for p, g in zip(params, grads):
                    p.grad = g.reshape(p.shape)
            optim.step()
            sched.step()
            if want_eval(s):
                torch.cuda.synchronize()
                te = time.perf_counter()
                evals[s] = psnr32([p.detach() for p in params], x, tgt)
                t_eval += time.perf_counter() - te
        losses = [float(loss)]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0 - t_eval
    tail = [v for s, v in sorted(evals.items()) if s > args.steps - args.tail]
    row = {"arm": arm, "image": idx, "steps": args.steps, "trailing_median_psnr": statistics.median(tail),
           "tail_min": min(tail), "tail_max": max(tail), "final_psnr": evals[args.steps],
           "curve": {str(s): round(evals[s], 3) for s in (100, 200, 500, 1000, 1500) if s in evals},
           "fit_seconds": dt, "steps_per_s": args.steps / dt, "final_loss": losses[-1]}
    return row


def check_emulation(args, dev):
    """emu_none must reproduce autograd's gradients (validates the explicit backward used by the emu arms)."""
    from implicit_image_compression_b200.data import get_grid, synth_image
    grid = get_grid(64, 96, dev)
    img = synth_image(64, 96, 0, device=dev).view(-1, 3)
    x = ((grid.view(-1, 2) - 0.5) * 2).contiguous()
    params = [p.requires_grad_(True) for p in init_params(args.depth, args.hidden, dev)]
    loss = torch.nn.functional.mse_loss(fwd32(params, x), img)
    loss.backward()
    with torch.no_grad():
        l2, grads = emu_loss_and_grads(params, x, img, EMU["emu_none"], {})
    worst = max(((g.reshape(p.shape) - p.grad).norm() / p.grad.norm()).item() for p, g in zip(params, grads))
    return {"check": "emu_none vs autograd", "loss_rel": abs(l2.item() - loss.item()) / loss.item(), "worst_grad_rel": worst}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", default="0-3")
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--tail", type=int, default=400)
    ap.add_argument("--eval-every", type=int, default=25)
    ap.add_argument("--arms", default="ref32,ref32_b,lib_f16tc")
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=768)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "psnr_study.jsonl"))
    args = ap.parse_args()
    a, _, b = args.images.partition("-")
    images = list(range(int(a), int(b or a) + 1))
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    rows = []
    with open(args.out, "a") as f:
        chk = check_emulation(args, dev)
        print(json.dumps(chk), flush=True)
        f.write(json.dumps(chk) + "\n")
        for idx in images:
            for arm in args.arms.split(","):
                row = run_arm(arm, idx, args, dev)
                rows.append(row)
                print(json.dumps(row), flush=True)
                f.write(json.dumps(row) + "\n")
                f.flush()
        # summary: mean trailing median per arm, and mean |delta| vs ref32 per image
        summary = {"summary": True, "images": images, "steps": args.steps, "tail": args.tail,
                   "eval_every": args.eval_every, "arms": {}}
        by = {}
        for r in rows:
            by.setdefault(r["arm"], {})[r["image"]] = r
        for arm, d in by.items():
            s = {"mean_trailing_median_psnr": statistics.mean(v["trailing_median_psnr"] for v in d.values()),
                 "mean_steps_per_s": statistics.mean(v["steps_per_s"] for v in d.values())}
            if "ref32" in by and arm != "ref32":
                common = [i for i in d if i in by["ref32"]]
                if common:
                    deltas = [d[i]["trailing_median_psnr"] - by["ref32"][i]["trailing_median_psnr"] for i in common]
                    s["mean_delta_vs_ref32"] = statistics.mean(deltas)
                    s["mean_abs_delta_vs_ref32"] = statistics.mean(abs(v) for v in deltas)
            summary["arms"][arm] = s
        print(json.dumps(summary), flush=True)
        f.write(json.dumps(summary) + "\n")


if __name__ == "__main__":
    main()
