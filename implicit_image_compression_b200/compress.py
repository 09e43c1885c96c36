"""Fit driver mirroring the reference entry point (implicit_image/compress.py:54-269) for the hot path:
build image / grid / model, main fit loop (:137-170), quantisation fine-tune loop (:172-216).
Saving weights and entropy coding (:242-263) stay on the host and are out of scope, as are W&B and hydra
(`config.load_config` reads the same keys).  Usage:
    python -m implicit_image_compression_b200.compress mlp.hidden_size=256 mlp.depth=6 masking=none quant=none
"""
import logging
import sys
from copy import deepcopy

import torch

from .config import load_config
from .data import get_grid, synth_image
from .fit import Fitter
from .models import registry as model_registry
from .pipeline.quant import context as quant_context
from .utils.train_helper import eval_epoch, get_device, get_optimizer_lr_scheduler, setup_mask, train_epoch


def main(cfg, fast=True):
    """Returns a dict of the numbers the reference logs (PSNR, Quant PSNR, Density, ...)."""
    torch.manual_seed(cfg.seed)
    device = get_device(cfg.device)
    img = synth_image(cfg.img.height, cfg.img.width, cfg.img.get("index", 0), cfg.img.get("bits", 16))
    grid = get_grid(cfg.img.height, cfg.img.width)

    small_density = 1.0
    if cfg.get("masking") and cfg.masking.get("name") == "Small_Dense":
        small_density = cfg.masking.density
    model = model_registry[cfg.mlp.name](**cfg.mlp, small_dense_density=small_density)
    model, grid, img = model.to(device), grid.to(device), img.to(device)

    model.train()
    optim, lr_scheduler = get_optimizer_lr_scheduler(model, cfg.optim)
    mult = cfg.train.multiplier
    num_steps = cfg.train.num_steps * mult
    masking_cfg = cfg.get("masking") or None
    if masking_cfg:
        if masking_cfg.get("end_when"):
            masking_cfg["end_when"] = int(masking_cfg["end_when"] * mult)
        if masking_cfg.get("interval"):
            masking_cfg["interval"] = int(masking_cfg["interval"] * mult)
    mask = setup_mask(model, optim, masking_cfg)
    out = {}

    if fast:
        fitter = Fitter(model, optim, grid, img, lr_scheduler, mask, masking_cfg if mask else None)
        done = 0
        while done < num_steps:
            k = min(cfg.train.log_steps, num_steps - done)
            fitter.steps(k)
            done += k
            _, loss, psnr, psnr8 = eval_epoch(model, grid, img)
            out.update(loss=loss, PSNR=psnr, PSNR_8bit=psnr8)
            logging.info(f"Train | Step: {done} | loss: {loss:.6f} | PSNR: {psnr:.4f} | PSNR_8bit: {psnr8:.4f}")
    else:
        # the reference's loop verbatim (compress.py:131-143), GradScaler included: fp32 arithmetic, the scaler
        # only adds its skip-on-overflow behaviour (autocast is never entered)
        scaler = torch.amp.GradScaler("cuda") if (cfg.train.get("mixed_precision") and device.type == "cuda") else None
        for i in range(num_steps):
            train_epoch(model, optim, grid, img, lr_scheduler=lr_scheduler, mask=mask, scaler=scaler)
            if mask and i <= masking_cfg["end_when"] and i % masking_cfg["interval"] == 0:
                mask.update_connections()
            if (i + 1) % cfg.train.log_steps == 0:
                _, loss, psnr, psnr8 = eval_epoch(model, grid, img)
                out.update(loss=loss, PSNR=psnr, PSNR_8bit=psnr8)
    if mask:
        out.update({"Prune Rate": mask.prune_rate, "Density": mask.stats.total_density})

    if cfg.get("quant"):
        quantized_model = deepcopy(model)
        optim_q, sched_q = get_optimizer_lr_scheduler(quantized_model, cfg.optim, quantize_mode=True)
        quantized_model.train()
        # Reference quirk App. A.2 (compress.py:186-188 + train_helper.py:166-168): with a mask present the
        # reference passes the ORIGINAL model's Masking into the quant loop, so `mask.step()` steps the old
        # optimizer on the original model with its stale gradients and the quantised copy only changes through
        # re-clustering.  quant.replicate_reference_mask_bug=true reproduces exactly that (for a "Quant PSNR"
        # comparable with the reference); the default fine-tunes the quantised copy with the final masks
        # FROZEN on it, so pruned weights stay exactly zero and k-means keeps excluding them.
        replicate = bool(mask) and bool(dict(cfg.quant).get("replicate_reference_mask_bug", False))
        if mask and not replicate:
            q_params = dict(quantized_model.named_parameters())
            optim_q.fused_masks = {q_params[n]: mask.mask_dict[n].clone() for n, _ in mask._masked_parameters()}
        with quant_context.Quantize(quantized_model, optim_q, cfg.quant) as q:
            for i in range(cfg.quant.num_steps):
                if replicate:
                    train_epoch(quantized_model, optim_q, grid, img, lr_scheduler=sched_q, mask=mask)
                else:
                    train_epoch(quantized_model, optim_q, grid, img, lr_scheduler=sched_q)
                # compress.py:190-193: periodic evaluation INSIDE the context — its forward re-clusters the
                # weights, which is what leaves an exact-zero centroid behind for convert() (kmeans.py:73-98)
                if (i + 1) % cfg.quant.get("log_steps", 10) == 0:
                    _, loss, psnr, psnr8 = eval_epoch(quantized_model, grid, img)
                    logging.info(f"Quant | Step: {i + 1} | loss: {loss:.6f} | PSNR: {psnr:.4f}")
                    quantized_model.train()
        quantized_model = q.convert()
        _, loss, psnr, psnr8 = eval_epoch(quantized_model, grid, img)
        out.update({"Quant loss": loss, "Quant PSNR": psnr, "Quant PSNR 8bit": psnr8})
        if mask:
            nz = tot = 0
            q_params = dict(quantized_model.named_parameters())
            for n, _ in mask._masked_parameters():
                nz += int((q_params[n] != 0).sum())
                tot += q_params[n].numel()
            out["Quant Density"] = nz / max(tot, 1)  # measured on the quantised weights, not the mask
            out["quant_replicates_reference_mask_bug"] = replicate
    return out


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    print(main(load_config(sys.argv[1:])))
