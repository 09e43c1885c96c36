"""Developer tool: clock64 timeline of block 0 of the fused tail kernel."""
import ctypes, os, sys
os.environ["SIRENB200_TIMELINE"] = "1"
os.environ["SIRENB200_TAIL"] = "1"
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
n = 3 * 4 * 8 * 16 + 12 * 16
buf = (ctypes.c_int64 * n)()
_lib.check(eng.lib.sirenb200_debug_timeline(eng.handle, buf, n))
v = list(buf)[3 * 4 * 8 * 16:]
t0 = min(x for x in v if x > 0)
print("tile | E: wait_acc acc_ok epi1_done y_ok epi2_done da0_ok half0_stored da1_ok tile_done | MMA: gemm_first_kb gemm_commit")
for t in range(12):
    row = v[t * 16:t * 16 + 11]
    print(t, " ".join(f"{(x - t0) if x else -1:7d}" for x in row))
