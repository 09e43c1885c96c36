"""Developer tool: clock64 timeline of block 0 of the layer-0-generating forward GEMM."""
import ctypes, os, sys
os.environ["SIRENB200_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from implicit_image_compression_b200 import _lib
from implicit_image_compression_b200.data import get_grid, synth_image
from implicit_image_compression_b200.models import Siren
torch.manual_seed(0)
model = Siren(depth=6, hidden_size=256, first_omega_0=50, hidden_omega_0=30, precision="f16tc").cuda()
grid, img = get_grid(512, 768, "cuda"), synth_image(512, 768, 0, device="cuda")
eng = model.engine_for(grid)
grads = [torch.empty_like(p) for p in model.hot_parameters()]
for _ in range(3):
    eng.forward_backward(model.kernel_parameters(), img, grads)
torch.cuda.synchronize()
n = 6 * 8 * 8
buf = (ctypes.c_int64 * n)()
_lib.check(eng.lib.sirenb200_debug_timeline(eng.handle, buf, n))
v = list(buf)
t0 = min(x for x in v if x > 0)
names = ["G0", "G1", "G2", "G3", "MMA", "EPI"]
print("G: k0 before a_empty wait, k1 after, k2 computed, k3 fenced, k4 arrived | MMA: k0 before tm_empty, k1 after, "
      "k2-5 a_full kb0-3, k6 committed | EPI: k0 before tm_full, k1 after, k2 tile done")
for role in range(6):
    for t in range(8):
        row = v[(role * 8 + t) * 8:(role * 8 + t) * 8 + 8]
        print(names[role], t, " ".join(f"{(x - t0) if x else -1:7d}" for x in row))
