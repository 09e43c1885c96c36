"""Trailing-window PSNR over several images (Adam trajectories of this problem are chaotic, so
equal-step PSNR is compared as the median of the last 5 evaluations, averaged over images).
Usage: python tools/ab_psnr.py [n_images] [steps]   (kernel variants are selected by SIRENB200_* env vars)"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from run_configs import fit_curve  # noqa: E402

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
prec = os.environ.get("AB_PRECISION", "f16tc")
rows = []
for idx in range(n_img):
    c = fit_curve(256, 6, 512, 768, steps, prec, idx=idx)
    tail = [p for _, _, p in c[-5:]]
    rows.append({"image": idx, "trailing_median_psnr": statistics.median(tail), "final_psnr": c[-1][2],
                 "min_tail": min(tail), "max_tail": max(tail), "seconds": c[-1][1]})
env = {k: v for k, v in os.environ.items() if k.startswith("SIRENB200_")}
print(json.dumps({"variant": env or "default", "precision": prec, "steps": steps, "images": n_img,
                  "mean_trailing_median_psnr": statistics.mean(r["trailing_median_psnr"] for r in rows),
                  "rows": rows}))
