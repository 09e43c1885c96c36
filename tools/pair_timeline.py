"""Developer tool: clock64 timeline of block 0 of the forward pair kernel (eval forward: no tail kernel)."""
import ctypes, os, sys
os.environ["SIRENB200_TIMELINE"] = "1"
os.environ["SIRENB200_FWD_PAIR"] = "1"
os.environ["SIRENB200_LAST_TC"] = "0"  # the tensor-core last-layer kernel stamps the same slots
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from implicit_image_compression_b200 import _lib
from implicit_image_compression_b200.data import get_grid, synth_image
from implicit_image_compression_b200.models import Siren
torch.manual_seed(0)
model = Siren(depth=5, hidden_size=256, first_omega_0=50, hidden_omega_0=30, precision="f16tc").cuda()
grid, img = get_grid(512, 768, "cuda"), synth_image(512, 768, 0, device="cuda")
eng = model.engine_for(grid)
for _ in range(3):
    eng.forward(model.kernel_parameters())
torch.cuda.synchronize()
n = 3 * 4 * 8 * 16 + 12 * 16
buf = (ctypes.c_int64 * n)()
_lib.check(eng.lib.sirenb200_debug_timeline(eng.handle, buf, n))
v = list(buf)[3 * 4 * 8 * 16:]
t0 = min(x for x in v if x > 0)
print("tile | E: wait_acc1 acc1_ok epi1_done acc2_ok epi2_done | MMA: g1_kb0 g1_commit g2_kb0 g2_commit g1_kb2 g2_kb2")
for t in range(12):
    row = v[t * 16:t * 16 + 11]
    print(t, " ".join(f"{(x - t0) if x else -1:7d}" for x in row))
