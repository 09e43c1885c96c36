"""Host-side cost of one train_epoch() call at config 2 (cProfile over 300 calls).  Dev tool."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from implicit_image_compression_b200.data import get_grid, synth_image  # noqa: E402
from implicit_image_compression_b200.models import Siren  # noqa: E402
from implicit_image_compression_b200.utils.train_helper import get_optimizer_lr_scheduler, train_epoch  # noqa: E402

H, W = 512, 768
torch.manual_seed(0)
model = Siren(depth=6, hidden_size=256, first_omega_0=50, hidden_omega_0=30, precision="f16tc").cuda()
grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")
optim, sched = get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
for _ in range(5):
    train_epoch(model, optim, grid, img, lr_scheduler=sched)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(300):
    train_epoch(model, optim, grid, img, lr_scheduler=sched)
torch.cuda.synchronize()
print("ms per train_epoch call:", (time.perf_counter() - t) / 300 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    train_epoch(model, optim, grid, img, lr_scheduler=sched)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
