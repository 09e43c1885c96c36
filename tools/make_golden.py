"""Generates tests/golden/*.npz from the UNMODIFIED reference imported from /root/reference (build container
only; the GPU box has no reference).  Re-run with:  python tools/make_golden.py

Every array is produced by reference code (implicit_image.*) running on CPU fp32 with the installed torch;
nothing here calls the oracle restatement or the product package, except `synth_image` from the oracle as
the (reference-independent) input image generator.
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402
from siren_oracle import synth_image  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
ns = ref_import.load()
AD = ns.AttrDict

MLP = dict(name="siren", first_omega_0=50, hidden_omega_0=30, outermost_linear=True,
           simulate_quantization=False)


def build(seed, depth, hidden):
    torch.manual_seed(seed)
    return ns.model_registry["siren"](**dict(MLP, depth=depth, hidden_size=hidden), small_dense_density=1.0)


def t2n(t):
    return t.detach().cpu().numpy().copy()


def fit_case(tag, depth, hidden, H, W, steps, keep_traj=True):
    """forward / loss / grads at init, then `steps` reference train_epoch steps (Adam 3e-4 + StepLR)."""
    model = build(0, depth, hidden)
    grid = ns.get_grid(H, W)
    img = synth_image(H, W, 0)
    out = {"depth": depth, "hidden": hidden, "H": H, "W": W, "grid": t2n(grid), "img": t2n(img)}
    for i, p in enumerate(model.parameters()):
        out[f"param{i}"] = t2n(p)
    pred = model(grid)
    loss = torch.nn.functional.mse_loss(pred, img)
    loss.backward()
    out["pred"] = t2n(pred)
    out["loss"] = np.float32(loss.item())
    for i, p in enumerate(model.parameters()):
        out[f"grad{i}"] = t2n(p.grad)
    optim, sched = ns.get_optimizer_lr_scheduler(model, AD(name="adam", lr=3e-4))
    losses = []
    for _ in range(steps):
        losses.append(ns.train_epoch(model, optim, grid, img, lr_scheduler=sched))
    out["losses"] = np.array(losses, dtype=np.float64)
    if keep_traj:
        for i, p in enumerate(model.parameters()):
            out[f"param_after{i}"] = t2n(p)
    _, l, psnr, psnr8 = ns.eval_epoch(model, grid, img)
    out["eval"] = np.array([l, psnr, psnr8], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, f"fit_{tag}.npz"), **out)
    print(tag, "loss0", out["loss"], "loss_end", losses[-1], "psnr", psnr)


def masking_case(tag, masking_cfg, depth=4, hidden=32, H=16, W=16, steps=14):
    """Reference Masking driven by train_epoch; every update_connections call is recorded as
    (state before) -> (state after) so the product can replay single updates bit-exactly."""
    model = build(0, depth, hidden)
    grid = ns.get_grid(H, W)
    img = synth_image(H, W, 1)
    optim, sched = ns.get_optimizer_lr_scheduler(model, AD(name="adam", lr=3e-4))
    cfg = AD(masking_cfg)
    torch.manual_seed(123)  # the mask RNG stream starts here (test does the same)
    mask = ns.setup_mask(model, optim, cfg)
    names = list(mask.mask_dict.keys())
    out = {"names": np.array(names), "depth": depth, "hidden": hidden, "H": H, "W": W,
           "baseline_nonzero": mask.baseline_nonzero, "total_params": mask.total_params}
    for i, p in enumerate(model.parameters()):
        out[f"init_param{i}"] = t2n(p)
    for n in names:
        out[f"init_mask/{n}"] = t2n(mask.mask_dict[n])
    rates = []
    nupd = 0
    for step in range(steps):
        ns.train_epoch(model, optim, grid, img, lr_scheduler=sched, mask=mask)
        rates.append(mask.prune_rate)
        if step <= cfg.end_when and step % cfg.interval == 0:
            pre = f"upd{nupd}/"
            pnames = [n for n, _ in model.named_parameters()]
            for n, p in model.named_parameters():
                out[pre + "w_before/" + n] = t2n(p)
                out[pre + "g_before/" + n] = t2n(p.grad)
                st = optim.state[p]
                out[pre + "m_before/" + n] = t2n(st["exp_avg"])
                out[pre + "v_before/" + n] = t2n(st["exp_avg_sq"])
            for n in names:
                out[pre + "mask_before/" + n] = t2n(mask.mask_dict[n])
            out[pre + "scalars_before"] = np.array(
                [mask.prune_threshold, mask.prune_rate, mask.mask_step, mask.adjusted_growth,
                 mask.stats.total_nonzero, mask.stats.total_zero], dtype=np.float64)
            out[pre + "adjustments_before"] = np.array(mask.adjustments, dtype=np.float64)
            mask.update_connections()
            for n, p in model.named_parameters():
                out[pre + "w_after/" + n] = t2n(p)
            for n in names:
                out[pre + "mask_after/" + n] = t2n(mask.mask_dict[n])
            out[pre + "scalars_after"] = np.array(
                [mask.prune_threshold, mask.prune_rate, mask.mask_step, mask.adjusted_growth,
                 mask.stats.total_nonzero, mask.stats.total_zero], dtype=np.float64)
            nupd += 1
            _ = pnames
    out["num_updates"] = nupd
    out["prune_rates"] = np.array(rates, dtype=np.float64)
    out["final_density"] = mask.stats.total_density
    np.savez_compressed(os.path.join(OUT, f"masking_{tag}.npz"), **out)
    print(tag, "updates", nupd, "density", mask.stats.total_density)


def decay_case():
    out = {}
    d = ns.decay_registry["cosine"](prune_rate=0.1, T_max=1500)
    seq = [d.get_dr()]
    for s in range(0, 60):
        d.step(s)
        seq.append(d.get_dr())
    out["cosine"] = np.array(seq, dtype=np.float64)
    d = ns.decay_registry["magnitude-prune"](final_sparsity=0.9, T_max=1500, T_start=5, interval=10)
    seq = []
    for s in range(0, 200):
        d.step(s, 0.001 * s)
        seq.append(d.get_dr())
    out["magnitude_prune"] = np.array(seq, dtype=np.float64)
    d = ns.decay_registry["linear"](prune_rate=0.3, T_max=100)
    seq = []
    for s in range(0, 120):
        d.step(s)
        seq.append(d.get_dr())
    out["linear"] = np.array(seq, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "decay.npz"), **out)


def quant_case():
    out = {}
    torch.manual_seed(7)
    w = (torch.rand(64, 64) - 0.5) * 0.02
    w[torch.rand(64, 64) < 0.3] = 0
    out["w"] = t2n(w)
    for bits in (4, 8):
        lin = torch.nn.Linear(64, 64)
        lin.weight.data = w.clone()
        km = ns.KmeansQuant.__new__(ns.KmeansQuant)
        km.bits = bits
        c, l, nw = ns.KmeansQuant.find_centroids(km, lin)
        out[f"kmeans{bits}_centroids"] = t2n(c)
        out[f"kmeans{bits}_labels"] = t2n(l)
        out[f"kmeans{bits}_weight"] = t2n(nw)
        nz = w[w != 0]
        out[f"kmeans{bits}_init"] = t2n(torch.linspace(nz.min(), nz.max(), 2 ** bits - 1))
    # Quantize(KMeans) context end to end on a small model: codes after convert()
    model = build(0, 4, 32)
    grid = ns.get_grid(12, 12)
    img = synth_image(12, 12, 2)
    optim, sched = ns.get_optimizer_lr_scheduler(model, AD(name="adam", lr=3e-4), quantize_mode=True)
    for i, p in enumerate(model.parameters()):
        out[f"q_param{i}"] = t2n(p)
    qcfg = AD(name="KMeans", bits=4, skip_ll=["layers.0.linear", "layers.3.linear"], num_steps=3)
    with ns.Quantize(model, optim, qcfg) as q:
        losses = [ns.train_epoch(model, optim, grid, img, lr_scheduler=sched) for _ in range(3)]
    qm = q.convert()
    out["q_losses"] = np.array(losses, dtype=np.float64)
    for name, module in qm.named_modules():
        if hasattr(module, "labeled_weight"):
            out[f"q_labels/{name}"] = t2n(module.labeled_weight)
            out[f"q_centroids/{name}"] = t2n(module.centroids)
    # QAT weight fake quant of the installed torch
    torch.manual_seed(8)
    w2 = torch.randn(48, 40) * 0.01
    fq = torch.quantization.get_default_qat_qconfig("fbgemm").weight()
    deq = fq(w2)
    out["fq_w"] = t2n(w2)
    out["fq_deq"] = t2n(deq)
    out["fq_scale"] = t2n(fq.scale)
    out["fq_codes"] = t2n(torch.round(deq / fq.scale[:, None])).astype(np.int8)
    np.savez_compressed(os.path.join(OUT, "quant.npz"), **out)
    print("quant ok")


def fourier_case():
    """FourierNet (models/fourier.py) of the reference: prediction, loss, gradients at init and a short Adam fit."""
    torch.manual_seed(0)
    model = ns.model_registry["fourier"](name="fourier", depth=5, hidden_size=48, map_size=32, map_scale=16,
                                         small_dense_density=1.0)
    H, W = 20, 28
    grid = ns.get_grid(H, W)
    img = synth_image(H, W, 1)
    out = {"depth": 5, "hidden": 48, "map_size": 32, "map_scale": 16, "grid": t2n(grid), "img": t2n(img)}
    names = [n for n, _ in model.named_parameters()]
    out["names"] = np.array(names)
    for n, p in model.named_parameters():
        out["param/" + n] = t2n(p)
    pred = model(grid)
    loss = torch.nn.functional.mse_loss(pred, img)
    loss.backward()
    out["pred"], out["loss"] = t2n(pred), np.float32(loss.item())
    for n, p in model.named_parameters():
        if p.grad is not None:
            out["grad/" + n] = t2n(p.grad)
    optim, sched = ns.get_optimizer_lr_scheduler(model, AD(name="adam", lr=3e-4))
    losses = [ns.train_epoch(model, optim, grid, img, lr_scheduler=sched) for _ in range(10)]
    out["losses"] = np.array(losses, dtype=np.float64)
    _, l_, psnr, _ = ns.eval_epoch(model, grid, img)
    out["eval_loss"], out["eval_psnr"] = np.float64(l_), np.float64(psnr)
    np.savez_compressed(os.path.join(OUT, "fourier.npz"), **out)
    print("fourier ok", losses[:3])


def entropy_case():
    """compress_state_dict / decompress_state_dict of the reference (pipeline/entropy_coding/__init__.py) with the
    'plain' stream on a k-means-quantised model: the exact bytes of compressed_weights.data and the decoded tensors."""
    import json
    import tempfile
    ec = ref_import.importlib_import("implicit_image.pipeline.entropy_coding")
    model = build(0, 4, 32)
    grid = ns.get_grid(12, 12)
    img = synth_image(12, 12, 2)
    optim, sched = ns.get_optimizer_lr_scheduler(model, AD(name="adam", lr=3e-4), quantize_mode=True)
    qcfg = AD(name="KMeans", bits=4, skip_ll=["layers.0.linear", "layers.3.linear"], num_steps=2)
    out = {}
    with ns.Quantize(model, optim, qcfg) as q:
        for _ in range(2):
            ns.train_epoch(model, optim, grid, img, lr_scheduler=sched)
        ns.eval_epoch(model, grid, img)
    qm = q.convert()
    for k, v in qm.state_dict().items():
        out["sd/" + k] = t2n(v)
    with tempfile.TemporaryDirectory() as d:
        size = ec.compress_state_dict(qm.half(), d, stream_name="plain")
        out["bytes"] = np.frombuffer(open(os.path.join(d, "compressed_weights.data"), "rb").read(), dtype=np.uint8)
        out["meta_json"] = np.array(open(os.path.join(d, "meta_data.json")).read())
        assert size == out["bytes"].size
        dec = ec.decompress_state_dict(d, stream_name="plain")
    for k, v in dec.items():
        out["dec/" + k] = t2n(v)
    np.savez_compressed(os.path.join(OUT, "entropy.npz"), **out)
    print("entropy ok", size, "bytes")


def qat_case():
    """Quantize(QAT) of the reference on its own Siren (quant/context.py:28-47): prepare_qat, 6 train_epoch steps
    with an eval_epoch in between, convert() -> per-step losses, activation observers, int8 weights."""
    out = {}
    model = build(0, 4, 32)
    grid = ns.get_grid(12, 12)
    img = synth_image(12, 12, 2)
    optim, sched = ns.get_optimizer_lr_scheduler(model, AD(name="adam", lr=3e-4), quantize_mode=True)
    for i, p in enumerate(model.parameters()):
        out[f"param{i}"] = t2n(p)
    out["grid"], out["img"] = t2n(grid), t2n(img)
    model.train()  # prepare_qat asserts training mode on current torch
    qcfg = AD(name="QAT", qconfig="fbgemm", num_steps=6)
    losses, evals = [], []
    with ns.Quantize(model, optim, qcfg) as q:
        for i in range(6):
            losses.append(ns.train_epoch(model, optim, grid, img, lr_scheduler=sched))
            if i == 2:
                _, l_, p_, _ = ns.eval_epoch(model, grid, img)
                evals.append(l_)
        for i, layer in enumerate(model.layers):
            app = layer.linear.activation_post_process
            out[f"act_min{i}"] = t2n(app.activation_post_process.min_val)
            out[f"act_max{i}"] = t2n(app.activation_post_process.max_val)
            out[f"act_scale{i}"] = t2n(app.scale)
            out[f"act_zp{i}"] = t2n(app.zero_point)
            out[f"w_scale{i}"] = t2n(layer.linear.weight_fake_quant.scale)
        for i, p in enumerate(model.parameters()):
            out[f"param_after{i}"] = t2n(p)
    out["losses"] = np.array(losses, dtype=np.float64)
    out["eval_losses"] = np.array(evals, dtype=np.float64)
    qm = q.convert()
    for i, layer in enumerate(qm.layers):
        lin = layer.linear
        out[f"int8_w{i}"] = t2n(lin.weight().int_repr())
        out[f"int8_w_scale{i}"] = t2n(lin.weight().q_per_channel_scales()).astype(np.float32)
        out[f"int8_out_scale{i}"] = np.float32(lin.scale)
        out[f"int8_out_zp{i}"] = np.int32(lin.zero_point)
    np.savez_compressed(os.path.join(OUT, "qat.npz"), **out)
    print("qat ok", losses)


if __name__ == "__main__":
    if "--only-qat" in sys.argv:
        qat_case()
        sys.exit(0)
    if "--only-entropy" in sys.argv:
        entropy_case()
        sys.exit(0)
    if "--only-fourier" in sys.argv:
        fourier_case()
        sys.exit(0)
    fit_case("d3_w16", 3, 16, 8, 10, 20)
    fit_case("d4_w128", 4, 128, 24, 32, 10)
    fit_case("d3_w256", 3, 256, 16, 24, 3, keep_traj=False)
    fit_case("d3_w64_c1small", 3, 64, 32, 32, 60, keep_traj=False)
    masking_case("pruning", dict(name="Pruning", density=1.0, sparse_init="random", final_density=0.5,
                                 dense_gradients=True, growth_mode="none", prune_mode="global-magnitude",
                                 redistribution_mode="none", dense=False,
                                 decay_schedule="magnitude-prune", start_when=2, end_when=12, interval=3))
    masking_case("rigl", dict(name="RigL", density=0.5, sparse_init="erdos-renyi-kernel",
                              dense_gradients=True, growth_mode="absolute-gradient",
                              prune_mode="magnitude", redistribution_mode="none", dense=False,
                              prune_rate=0.1, decay_schedule="cosine", end_when=12, interval=4))
    masking_case("snfs", dict(name="SNFS", density=0.3, sparse_init="erdos-renyi-kernel",
                              dense_gradients=True, growth_mode="momentum", prune_mode="magnitude",
                              redistribution_mode="momentum", dense=False, prune_rate=0.1,
                              decay_schedule="cosine", end_when=12, interval=4))
    decay_case()
    quant_case()
    qat_case()
    fourier_case()
    entropy_case()
    copy  # noqa
    print("golden written to", OUT)
