"""Per-kernel device time of sirenb200_kmeans_quantize on a few weight distributions.  Dev tool."""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from implicit_image_compression_b200 import engine as eng  # noqa: E402

torch.manual_seed(0)
cases = {}
w = torch.randn(256, 256, device="cuda") * 0.02
cases["randn256"] = w
wp = w.clone()
wp[wp.abs() < 0.02] = 0
cases["pruned256"] = wp
cases["quantised256"] = eng.kmeans_quantize(wp, 8)[2]
cases["randn512"] = torch.randn(512, 512, device="cuda") * 0.02
for name, x in cases.items():
    eng.kmeans_quantize(x, 8)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        eng.kmeans_quantize(x, 8)
    torch.cuda.synchronize()
    print(name, "ms per call", (time.perf_counter() - t) / 10 * 1e3)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.kmeans_quantize(x, 8)
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if e.device_time_total > 0:
            print("   ", e.key[:60], e.count, "x", round(e.device_time_total / max(e.count, 1), 1), "us")
