"""Developer tool: where the single-thread roles of the GEMM-class kernels wait (SIRENB200_STALLS=1).
Per launch slot (forward layer l = slot l, backward layer l = slot 6 + l) and per role, mean cycles over CTAs:
  row GEMM  loader: pace, a_empty, total | MMA: tm_empty, a_full, total | E loader: eo_empty |
            epilogue (thread 128): tm_full, eo_full/eo_empty, named barrier, store wait_read, total
  reduction loader: pace, empty, total | MMA: full, total
"""
import ctypes, os, sys
os.environ["SIRENB200_STALLS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from implicit_image_compression_b200 import _lib
from implicit_image_compression_b200.data import get_grid, synth_image
from implicit_image_compression_b200.models import Siren

hidden = int(os.environ.get("HIDDEN", 256))
depth = int(os.environ.get("DEPTH", 6))
H, W = int(os.environ.get("IMG_H", 512)), int(os.environ.get("IMG_W", 768))
torch.manual_seed(0)
model = Siren(depth=depth, hidden_size=hidden, first_omega_0=50, hidden_omega_0=30, precision="f16tc").cuda()
grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")
eng = model.engine_for(grid)
grads = [torch.empty_like(p) for p in model.hot_parameters()]
nsteps = int(os.environ.get("STEPS", 4))
for _ in range(3):
    eng.forward_backward(model.kernel_parameters(), img, grads)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(nsteps):
    eng.forward_backward(model.kernel_parameters(), img, grads)
e1.record()
torch.cuda.synchronize()
print(f"{nsteps} eager forward_backward calls: {e0.elapsed_time(e1) / nsteps * 1e3:.1f} us per call (no Adam, no graph)")
base = 3 * 4 * 8 * 16 + 12 * 16
n = base + 12 * 160 * 16
buf = (ctypes.c_int64 * n)()
_lib.check(eng.lib.sirenb200_debug_timeline(eng.handle, buf, n))
v = list(buf)[base:]
names_row = ["ld.pace", "ld.a_empty", "ld.total", "mma.tm_empty", "mma.a_full", "mma.total", "eld.eo_empty",
             "epi.tm_full", "epi.eo", "epi.bar", "epi.store_rd", "epi.total"]
names_col = ["ld.pace", "ld.empty", "ld.total", "-", "mma.full", "mma.total"]
for slot in range(12):
    rows = [v[(slot * 160 + c) * 16:(slot * 160 + c) * 16 + 16] for c in range(160)]
    row_ctas = [r for r in rows if r[11] > 0 or r[5] > 0 and r[2] == 0]
    rowg = [r for r in rows if r[11] > 0]
    colg = [r for r in rows if r[11] == 0 and r[5] > 0]
    if not rowg and not colg:
        continue
    print(f"--- slot {slot} ({'forward' if slot < 6 else 'backward'} layer {slot % 6}): {len(rowg)} row-GEMM CTAs, "
          f"{len(colg)} reduction CTAs")
    if rowg:
        print("  row : " + "  ".join(f"{nm}={sum(r[k] for r in rowg) / len(rowg):.0f}" for k, nm in enumerate(names_row)))
        print("  row max: " + "  ".join(f"{nm}={max(r[k] for r in rowg)}" for k, nm in enumerate(names_row)))
    if rowg:
        life_c = sum(r[12] for r in rowg) / len(rowg)
        life_ns = sum(r[13] for r in rowg) / len(rowg)
        t0, t1 = min(r[14] for r in rowg), max(r[15] for r in rowg)
        if life_ns > 0:
            print(f"  row CTA lifetime: {life_c:.0f} cycles = {life_ns / 1e3:.1f} us -> SM clock {life_c / life_ns * 1e3:.0f} MHz; "
                  f"launch span (first CTA start .. last CTA end) {(t1 - t0) / 1e3:.1f} us; first start @{t0 % 10**9 / 1e3:.1f} us, "
                  f"last end @{t1 % 10**9 / 1e3:.1f} us")
    if colg:
        print("  red : " + "  ".join(f"{nm}={sum(r[k] for r in colg) / len(colg):.0f}" for k, nm in enumerate(names_col) if nm != "-"))
        print("  red max: " + "  ".join(f"{nm}={max(r[k] for r in colg)}" for k, nm in enumerate(names_col) if nm != "-"))
