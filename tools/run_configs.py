"""Runs the BASELINE.json configurations end to end on one B200 and prints one JSON line per result:
   c1  W64 D3 256x256 (fp32 CUDA-core path), c2 W256 D6 512x768 (time-to-PSNR vs the fp32 path),
   c4  c2 shape + Pruning (global-magnitude, final density 0.1) + 8-bit k-means fine-tune,
   c5  {256,512}x{6,8} sweep on one image each -> images/hour/GPU.
Usage: python tools/run_configs.py [c1] [c2] [c4] [c5] [--steps N]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from implicit_image_compression_b200.config import load_config  # noqa: E402
from implicit_image_compression_b200 import compress  # noqa: E402
from implicit_image_compression_b200.data import get_grid, synth_image  # noqa: E402
from implicit_image_compression_b200.fit import Fitter  # noqa: E402
from implicit_image_compression_b200.models import Siren  # noqa: E402
from implicit_image_compression_b200.utils.train_helper import eval_epoch, get_optimizer_lr_scheduler  # noqa: E402


def fit_curve(hidden, depth, H, W, steps, precision, idx=0, every=100):
    """Returns [(step, seconds_of_fitting, psnr)], eval time excluded (SURVEY.md §8d time-to-PSNR)."""
    torch.manual_seed(0)
    model = Siren(depth=depth, hidden_size=hidden, first_omega_0=50, hidden_omega_0=30,
                  precision=precision).cuda()
    grid, img = get_grid(H, W, "cuda"), synth_image(H, W, idx, device="cuda")
    optim, sched = get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
    fitter = Fitter(model, optim, grid, img, sched)
    fitter.steps(3)  # warm-up (counted in the step index, not in the clock)
    torch.cuda.synchronize()
    out, spent, done = [], 0.0, 3
    while done < steps:
        k = min(every - done % every, steps - done)
        t = time.perf_counter()
        fitter.steps(k)
        torch.cuda.synchronize()
        spent += time.perf_counter() - t
        done += k
        out.append((done, spent, eval_epoch(model, grid, img)[2]))
    return out


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c2", "c4", "c5"]
    steps = 2000
    if "--steps" in sys.argv:
        steps = int(sys.argv[sys.argv.index("--steps") + 1])
    if "c1" in which:
        c = fit_curve(64, 3, 256, 256, steps, "fp32")
        print(json.dumps({"config": "c1 W64 D3 256x256 fp32 path", "steps": steps, "seconds": c[-1][1],
                          "steps_per_s": (steps - 3) / c[-1][1], "psnr_curve": {s: round(p, 3) for s, _, p in c
                                                                                if s in (100, 200, 500, 1000, 2000)},
                          "oracle_psnr_2000_reference_cpu": 32.02}))
    if "c2" in which:
        ref = fit_curve(256, 6, 512, 768, steps, "fp32")
        tc = fit_curve(256, 6, 512, 768, steps, "f16tc")
        target = ref[-1][2]
        reach = next(((s, t) for s, t, p in tc if p >= target - 0.1), None)
        marks = (100, 200, 300, 500, 1000, 1500, 2000)
        print(json.dumps({
            "config": "c2 W256 D6 512x768", "steps": steps,
            "target_psnr_fp32_path": target,
            "f16tc_time_to_target_minus_0.1dB_s": None if reach is None else reach[1],
            "f16tc_steps_to_target": None if reach is None else reach[0],
            "f16tc_total_s": tc[-1][1], "fp32_total_s": ref[-1][1],
            "psnr_f16tc": {s: round(p, 3) for s, _, p in tc if s in marks},
            "psnr_fp32": {s: round(p, 3) for s, _, p in ref if s in marks},
            "trailing_median_psnr_f16tc": sorted(p for _, _, p in tc[-5:])[2],
            "trailing_median_psnr_fp32": sorted(p for _, _, p in ref[-5:])[2]}))
    if "c4" in which:
        cfg = load_config(["mlp.hidden_size=256", "mlp.depth=6", "img.height=512", "img.width=768",
                           "masking=Pruning", "masking.final_density=0.1", "quant=kmeans", "quant.bits=8",
                           f"train.num_steps={steps}", f"train.log_steps={steps}",
                           f"masking.end_when={int(0.75 * steps)}"])
        cfg.quant["skip_ll"] = ["layers.0.linear", "layers.5.linear"]
        t = time.perf_counter()
        res = compress.main(cfg)
        torch.cuda.synchronize()
        res["seconds"] = time.perf_counter() - t
        res["config"] = "c4 W256 D6 512x768 Pruning(global-magnitude -> density 0.1) + KMeans 8 bit (100 steps)"
        print(json.dumps(res))
    if "c5" in which:
        rows = []
        for hidden in (256, 512):
            for depth in (6, 8):
                c = fit_curve(hidden, depth, 512, 768, steps, "f16tc", idx=1)
                rows.append({"hidden": hidden, "depth": depth, "seconds": c[-1][1], "psnr": round(c[-1][2], 3),
                             "steps_per_s": (steps - 3) / c[-1][1]})
        per_image = sum(r["seconds"] for r in rows) / len(rows)
        print(json.dumps({"config": f"c5 sweep {{256,512}}x{{6,8}}, {steps} steps per fit, 512x768", "fits": rows,
                          "images_per_hour_per_gpu": 3600.0 / per_image,
                          "note": "one independent fit per GPU, no communication: 8 GPUs = 8x"}))


if __name__ == "__main__":
    main()
