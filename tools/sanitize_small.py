"""Small end-to-end exercise of every kernel family, the target of `compute-sanitizer --tool memcheck`."""
import os
import sys

os.environ.setdefault("SIRENB200_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from implicit_image_compression_b200 import engine as eng  # noqa: E402
from implicit_image_compression_b200.data import get_grid, synth_image  # noqa: E402
from implicit_image_compression_b200.models import Siren  # noqa: E402
from implicit_image_compression_b200.utils.train_helper import (eval_epoch, get_optimizer_lr_scheduler,  # noqa: E402
                                                                 train_epoch)

for hidden, depth, H, W, prec in ((256, 4, 37, 53, "f16tc"), (128, 3, 16, 24, "f16tc"), (512, 3, 24, 32, "f16tc"),
                                  (48, 3, 9, 11, "fp32")):
    torch.manual_seed(0)
    model = Siren(depth=depth, hidden_size=hidden, first_omega_0=50, hidden_omega_0=30, precision=prec).cuda()
    grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")
    optim, sched = get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
    losses = [train_epoch(model, optim, grid, img, lr_scheduler=sched) for _ in range(3)]
    psnr = eval_epoch(model, grid, img)[2]
    print(hidden, depth, H, W, prec, [round(x, 5) for x in losses], round(psnr, 2))
w = torch.randn(128, 128, device="cuda") * 0.02
w[w.abs() < 0.01] = 0
c, l, q = eng.kmeans_quantize(w, 5)
codes, scales, wq = eng.fakequant_per_channel(w)
torch.cuda.synchronize()
print("kmeans centroids", c.numel(), "fakequant ok", codes.dtype)
