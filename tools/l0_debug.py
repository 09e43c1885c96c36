"""Dev tool: layer-0 gradient from the dX-fused reducer vs the stand-alone kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from implicit_image_compression_b200.data import get_grid, synth_image
from implicit_image_compression_b200.models import Siren

H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 96
grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")
out = {}
for mode in ("1", "0"):
    os.environ["SIRENB200_FUSE_L0"] = mode
    torch.manual_seed(0)
    model = Siren(depth=4, hidden_size=256, first_omega_0=50, hidden_omega_0=30, precision="f16tc").cuda()
    eng = model.engine_for(grid)
    grads = [torch.zeros_like(p) for p in model.hot_parameters()]
    for _ in range(2):
        eng.forward_backward(model.kernel_parameters(), img, grads)
    torch.cuda.synchronize()
    out[mode] = [g.clone() for g in grads]
a, b = out["1"], out["0"]
for i, (x, y) in enumerate(zip(a, b)):
    print(i, tuple(x.shape), "rel", ((x - y).norm() / y.norm()).item())
dw_f, dw_r = a[0], b[0]
print("ratio dW0[:8]:", (dw_f[:8] / dw_r[:8]).tolist())
print("db0 ratio[:8]:", (a[1][:8] / b[1][:8]).tolist())
print("db0 ratio[64:72]:", (a[1][64:72] / b[1][64:72]).tolist())
