"""Dev tool: where the per-call gap of train_epoch (graph replay + loss read) comes from."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_image_compression_b200.data import get_grid, synth_image
from implicit_image_compression_b200.fit import Fitter
from implicit_image_compression_b200.models import Siren
from implicit_image_compression_b200.utils.train_helper import get_optimizer_lr_scheduler

H, W = 512, 768
torch.manual_seed(0)
model = Siren(depth=6, hidden_size=256, first_omega_0=50, hidden_omega_0=30, precision="f16tc").cuda()
grid, img = get_grid(H, W, "cuda"), synth_image(H, W, 0, device="cuda")
optim, sched = get_optimizer_lr_scheduler(model, {"name": "adam", "lr": 3e-4})
f = Fitter(model, optim, grid, img, sched)
f.steps(5)
torch.cuda.synchronize()
g = f._graph
st = torch.cuda.current_stream()


def clock(fn, n=300):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


def spin():
    g.replay()
    while not st.query():
        pass


ev = torch.cuda.Event()


def ev_sync():
    g.replay()
    ev.record()
    ev.synchronize()


print("back-to-back replays      ms/step", clock(g.replay))
print("replay + stream.sync      ms/step", clock(lambda: (g.replay(), st.synchronize())))
print("replay + query spin       ms/step", clock(spin))
print("replay + event.sync       ms/step", clock(ev_sync))
print("2 replays + stream.sync   ms/step", clock(lambda: (g.replay(), g.replay(), st.synchronize())) / 2)
print("Fitter.step_loss          ms/step", clock(f.step_loss))
