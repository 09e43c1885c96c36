"""Where config 4 spends its time: masked fit segments, update_connections, k-means transform, quant-phase
step.  Dev tool; prints one JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from implicit_image_compression_b200 import engine as eng  # noqa: E402
from implicit_image_compression_b200.config import load_config  # noqa: E402
from implicit_image_compression_b200.data import get_grid, synth_image  # noqa: E402
from implicit_image_compression_b200.fit import Fitter  # noqa: E402
from implicit_image_compression_b200.models import Siren  # noqa: E402
from implicit_image_compression_b200.pipeline.quant import context as quant_context  # noqa: E402
from implicit_image_compression_b200.utils.train_helper import (eval_epoch, get_optimizer_lr_scheduler,  # noqa: E402
                                                                 setup_mask, train_epoch)


def clock(fn, n=1):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n


H, W = 512, 768
out = {}
torch.manual_seed(0)
cfg = load_config(["mlp.hidden_size=256", "mlp.depth=6", "masking=Pruning", "masking.final_density=0.1",
                   "quant=kmeans", "quant.bits=8"])
cfg.quant["skip_ll"] = ["layers.0.linear", "layers.5.linear"]
model = Siren(depth=6, hidden_size=256, first_omega_0=50, hidden_omega_0=30).cuda()
grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")
optim, sched = get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
mcfg = dict(cfg.masking)
mask = setup_mask(model, optim, mcfg)
f = Fitter(model, optim, grid, img, sched, mask, mcfg)
f.steps(30)
out["masked_fit_ms_per_step_incl_updates"] = clock(lambda: f.steps(200)) / 200 * 1e3
out["update_connections_ms"] = clock(mask.update_connections, 5) * 1e3
out["density"] = mask.stats.total_density
w = model.layers[2].linear.weight.detach()
out["kmeans_quantize_ms_W256"] = clock(lambda: eng.kmeans_quantize(w, 8), 10) * 1e3
w5 = torch.randn(512, 512, device="cuda") * 0.01
out["kmeans_quantize_ms_W512"] = clock(lambda: eng.kmeans_quantize(w5, 8), 10) * 1e3
out["eval_epoch_ms"] = clock(lambda: eval_epoch(model, grid, img), 5) * 1e3
from copy import deepcopy  # noqa: E402
qm = deepcopy(model)
oq, sq = get_optimizer_lr_scheduler(qm, cfg.optim, quantize_mode=True)
qm.train()
with quant_context.Quantize(qm, oq, cfg.quant) as q:
    train_epoch(qm, oq, grid, img, lr_scheduler=sq)
    out["quant_phase_train_epoch_ms"] = clock(lambda: train_epoch(qm, oq, grid, img, lr_scheduler=sq), 10) * 1e3
print(json.dumps(out))
if "--profile" in sys.argv:
    import cProfile
    import pstats
    with quant_context.Quantize(deepcopy(model).train(), oq, cfg.quant):
        pass
    qm2 = deepcopy(model)
    oq2, sq2 = get_optimizer_lr_scheduler(qm2, cfg.optim, quantize_mode=True)
    qm2.train()
    with quant_context.Quantize(qm2, oq2, cfg.quant):
        train_epoch(qm2, oq2, grid, img, lr_scheduler=sq2)
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(5):
            train_epoch(qm2, oq2, grid, img, lr_scheduler=sq2)
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(14)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        mask.update_connections()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
