"""Runs a few c2 fit steps (the bench workload) — the target command for ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from implicit_image_compression_b200.data import get_grid, synth_image  # noqa: E402
from implicit_image_compression_b200.fit import Fitter  # noqa: E402
from implicit_image_compression_b200.models import Siren  # noqa: E402
from implicit_image_compression_b200.utils.train_helper import get_optimizer_lr_scheduler  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
H, W = 512, 768
torch.manual_seed(0)
model = Siren(depth=6, hidden_size=256, first_omega_0=50, hidden_omega_0=30, precision="f16tc").cuda()
grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")
optim, sched = get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
losses = Fitter(model, optim, grid, img, sched).steps(steps)
torch.cuda.synchronize()
print("losses", [round(x, 6) for x in losses.tolist()])
