"""Dev tool: a fused kernel variant (env switch given as argv[1], default SIRENB200_TAIL) on vs off: gradients, loss, time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
VAR = sys.argv[1] if len(sys.argv) > 1 else "SIRENB200_TAIL"
from implicit_image_compression_b200.data import get_grid, synth_image
from implicit_image_compression_b200.models import Siren

for hidden, depth, H, W in ((256, 4, 64, 96), (128, 5, 40, 56), (256, 6, 512, 768)):
    grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")
    out = {}
    for mode in ("1", "0"):
        os.environ[VAR] = mode
        torch.manual_seed(0)
        model = Siren(depth=depth, hidden_size=hidden, first_omega_0=50, hidden_omega_0=30, precision="f16tc").cuda()
        eng = model.engine_for(grid)
        grads = [torch.zeros_like(p) for p in model.hot_parameters()]
        stats = torch.zeros(4, device="cuda")
        for _ in range(3):
            eng.forward_backward(model.kernel_parameters(), img, grads, stats)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(20):
            eng.forward_backward(model.kernel_parameters(), img, grads, stats)
        torch.cuda.synchronize()
        out[mode] = ([g.clone() for g in grads], stats.clone(), (time.perf_counter() - t) / 20 * 1e3)
    a, b = out["1"], out["0"]
    rels = [((x - y).norm() / (y.norm() + 1e-30)).item() for x, y in zip(a[0], b[0])]
    print(f"W{hidden} D{depth} {H}x{W}: max grad rel diff {max(rels):.2e}  loss fused {a[1][1].item():.8f} "
          f"ref {b[1][1].item():.8f}  ms fused {a[2]:.3f} ref {b[2]:.3f}")
    print("   per-tensor:", [f"{r:.1e}" for r in rels])
