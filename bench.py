#!/usr/bin/env python
"""bench.py — SIREN fit steps/sec on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (all N): BASELINE.json configs[1] — SIREN width 256 / depth 6, full-image fit of a synthetic
512x768 16-bit RGB image; a "step" = forward + MSE + backward + fused Adam (+StepLR) over the whole image.
N > 1 (torchrun, one rank per GPU): the image rows are sharded over ranks, one NCCL all-reduce of the flat
gradient buffer per step, identical Adam on every rank (strong scaling of the same job).

Output: ONE JSON line (rank 0) with the contract keys plus `roofline`, `cpu_baseline`, `e2e`,
`gpu_launches`, `clocks`.  `--impl reference` times the reference's CPU path (oracle port — the reference
itself is Python and cannot travel to the GPU box) on the host cores instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# default workload = BASELINE.json configs[1]; --workload c3 (configs[2]: hidden 512, depth 8, 2048x2048) is a
# developer option for the pixel-sharded large-image case and is NOT what the driver's contract line measures
WORKLOADS = {"c2": (6, 256, 512, 768), "c3": (8, 512, 2048, 2048), "c5w": (6, 512, 512, 768)}  # c5w: the wide half of the sweep
DEPTH, HIDDEN, H, W, C = 6, 256, 512, 768, 3
OMEGA0, OMEGA = 50.0, 30.0
LR = 3e-4
METRIC, UNIT = "siren_fit_steps_per_sec", "steps/s"


def f_step(n_pix, depth=None, w=None):
    """Algorithmic FLOPs per fit step: 2 N [3 (D-2) W^2 + 13 W]."""
    depth = DEPTH if depth is None else depth
    w = HIDDEN if w is None else w
    return 2.0 * n_pix * (3 * (depth - 2) * w * w + 13 * w)


def workload_config(n_gpus):
    return {
        "workload": f"{'c2' if HIDDEN == 256 else 'c3'}: SIREN hidden {HIDDEN} depth {DEPTH}, {H}x{W} synthetic 16-bit RGB, dense Adam fit step"
                    + ("" if n_gpus == 1 else f", rows sharded over {n_gpus} ranks + one gradient all-reduce per step"),
        "hidden_size": HIDDEN, "depth": DEPTH, "height": H, "width": W, "pixels": H * W,
        "optimizer": "adam lr 3e-4 + StepLR(2000, 0.5)", "precision_mode": "f16tc",
        "l2": "no flush: each step streams ~2 GB of activation stash (>> 126 MB L2)",
        "algorithmic_flops_per_step": f_step(H * W),
    }


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.samples:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= ts <= t1 + 0.05:
                sm.append(clk)
                for nme, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
        if not sm:  # region shorter than one sample period: use every sample taken
            for ts, line in self.samples:
                parts = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(parts[0]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference's CPU path)
# ------------------------------------------------------------------------------------------------
def cpu_port_steps_per_sec(budget_s, steps_hint=None, warmup=1):
    """Times oracle forward+backward+Adam (the reference's train_epoch math) on the host cores.
    Returns (steps/s as full-image-equivalent, description, cores)."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import siren_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    params = O.siren_init(0, DEPTH, HIDDEN, OMEGA0, OMEGA)
    grid, img = O.get_grid(H, W), O.synth_image(H, W, 0)
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]

    def one_step(rows, step):
        loss, grads = O.siren_loss_and_grads(params, grid[:rows], img[:rows], OMEGA0, OMEGA)
        for i in range(len(params)):
            params[i], m[i], v[i] = O.adam_step(params[i], grads[i], m[i], v[i], step, LR)
        return loss

    # probe with a 32-row band to size the sample
    t = time.perf_counter()
    one_step(32, 1)
    t_band = time.perf_counter() - t
    est_full = t_band * (H / 32)
    n_steps = steps_hint if steps_hint else 3
    rows = H
    if est_full * (n_steps + warmup) > budget_s:
        rows = max(16, int(H * budget_s / (est_full * (n_steps + warmup))) // 16 * 16)
        rows = min(rows, H)
    for s in range(warmup):
        one_step(rows, 2 + s)
    t = time.perf_counter()
    for s in range(n_steps):
        one_step(rows, 2 + warmup + s)
    dt = time.perf_counter() - t
    value = n_steps * (rows / H) / dt
    desc = (f"{n_steps} oracle steps (fwd + MSE + bwd + Adam, torch CPU fp32) over a {rows}-row band of the "
            f"{H}x{W} image after {warmup} warm-up, scaled by {rows}/{H} to full-image steps")
    return value, desc, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, desc, cores = cpu_port_steps_per_sec(budget_s=150.0, steps_hint=max(1, args.steps),
                                                warmup=max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from implicit_image_compression_b200 import _lib
    from implicit_image_compression_b200.data import get_grid, synth_image
    from implicit_image_compression_b200.fit import Fitter
    from implicit_image_compression_b200.models import Siren
    from implicit_image_compression_b200.utils.train_helper import get_optimizer_lr_scheduler, train_epoch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    # NCCL prints its version banner on STDOUT at communicator creation: keep stdout clean for the one
    # JSON line by pointing fd 1 at stderr until the warm-up (first collective) is over
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    torch.manual_seed(0)
    model = Siren(depth=DEPTH, hidden_size=HIDDEN, first_omega_0=OMEGA0, hidden_omega_0=OMEGA,
                  precision="f16tc").to(dev)
    grid = get_grid(H, W, dev)
    img_host = synth_image(H, W, 0).pin_memory()
    img = img_host.to(dev, non_blocking=True)
    optim, sched = get_optimizer_lr_scheduler(model, {"name": "adam", "lr": LR})
    fitter = Fitter(model, optim, grid, img, sched, rank=rank, world_size=world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -----------------------------------------------------------
    fitter.steps(max(3, args.warmup))
    barrier()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    losses = fitter.steps(args.steps)
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    if fitter._graph is not None:
        # graph replays re-issue the launches captured once: count them per replayed step
        launches = args.steps * fitter.launches_per_step
    clocks = sampler.stop(t0, t1)
    # per-kernel device times: an immediately following EAGER pass of the same K steps with cudaEvent
    # pairs around every library kernel (events cannot be timed inside a replayed graph)
    graph_mode = fitter.use_graph
    fitter.use_graph = False
    fitter.engine.profile(True)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    pe0.record()
    fitter.steps(args.steps)
    pe1.record()
    barrier()
    ms_eager = pe0.elapsed_time(pe1)
    prof = fitter.engine.profile_read()
    fitter.engine.profile(False)
    fitter.use_graph = graph_mode
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = tms.item()
    ms_per_step = ms / args.steps
    value = 1000.0 / ms_per_step

    # ---- end to end through the drop-in API: pinned host image -> device every step, loss read back ----
    e2e = None
    if world == 1:
        copy_stream = torch.cuda.Stream()
        bufs = [torch.empty_like(img), torch.empty_like(img)]
        evs = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                bufs[i % 2].copy_(img_host, non_blocking=True)
                evs[i % 2].record(copy_stream)

        for b in bufs:
            b.copy_(img_host, non_blocking=True)
        for i in range(6):  # warm-up: train_epoch captures its step graph per image buffer on the 2nd call
            train_epoch(model, optim, grid, bufs[i % 2], lr_scheduler=sched)
        torch.cuda.synchronize()
        n_e2e = args.steps
        t = time.perf_counter()
        prefetch(0)
        for i in range(n_e2e):
            torch.cuda.current_stream().wait_event(evs[i % 2])
            if i + 1 < n_e2e:
                prefetch(i + 1)
            train_epoch(model, optim, grid, bufs[i % 2], lr_scheduler=sched)  # returns loss.item()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        e2e = {"value": n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": img_host.numel() * 4,
               "d2h_bytes_per_step": 4,
               "note": "train_epoch() drop-in API (one CUDA-graph replay per call); image copied from pinned "
                       "host memory every step (prefetched on a copy stream), loss.item() read back every step"}

    else:
        # pixel-sharded: every rank copies ITS rows of the image from pinned host memory each step and
        # rank 0 reads the (all-reduced) loss back each step
        shard_host = img_host[fitter.row_begin:fitter.row_end].contiguous().pin_memory()
        copy_stream = torch.cuda.Stream()
        stage = [torch.empty_like(fitter.img), torch.empty_like(fitter.img)]
        evs = [torch.cuda.Event(), torch.cuda.Event()]
        used = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(i):  # host -> device on the copy stream, overlapping the previous step
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(used[i % 2])  # the step that consumed this staging buffer is done
                stage[i % 2].copy_(shard_host, non_blocking=True)
                evs[i % 2].record(copy_stream)

        fitter.steps(3)
        barrier()
        n_e2e = args.steps
        t = time.perf_counter()
        prefetch(0)
        cur = torch.cuda.current_stream()
        for i in range(n_e2e):
            cur.wait_event(evs[i % 2])
            fitter.img.copy_(stage[i % 2], non_blocking=True)  # device -> device into the buffer the graph reads
            used[i % 2].record(cur)
            if i + 1 < n_e2e:
                prefetch(i + 1)
            _ = fitter.step_loss()  # one graph replay; the loss comes back through pinned host memory
        barrier()
        dt = torch.tensor([time.perf_counter() - t], device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_e2e / dt.item(), "unit": UNIT, "h2d_bytes_per_step": img_host.numel() * 4,
               "d2h_bytes_per_step": 4 * world,
               "note": "Fitter.step_loss() per step on every rank; each rank copies its image rows from pinned "
                       "host memory every step (prefetched on a copy stream into a staging buffer) and reads the "
                       "all-reduced loss back every step (max over ranks)"}

    def finish():
        # Tear-down: drop the captured graph (it holds NCCL kernels) before leaving, and leave without
        # running NCCL's destructor chain — destroy_process_group() after a captured collective can hang.
        sys.stdout.flush()
        if world > 1:
            fitter._graph = None
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        finish()
        return

    peaks = load_peaks()
    npix_rank = (fitter.row_end - fitter.row_begin) * W
    # dominant kernel = the tagged kind with the largest device time in the timed region
    kinds = {k: v for k, v in prof.items() if v[1] > 0}
    dom = max(kinds, key=lambda k: kinds[k][0]) if kinds else None
    alg_bytes = {  # algorithmic HBM bytes per launch (fp16 activations)
        "fwd_gemm": 2 * npix_rank * HIDDEN * 2,
        "dx_gemm": 3 * npix_rank * HIDDEN * 2,
        "dw_gemm": (DEPTH - 2) * 2 * npix_rank * HIDDEN * 2,
        "last_layer_loss": 2 * npix_rank * HIDDEN * 2 + npix_rank * C * 4,
        "first_layer": npix_rank * HIDDEN * 2,
        "layer0_grad": npix_rank * HIDDEN * 2,
    }
    roofline = None
    if dom in alg_bytes:
        avg_ms = kinds[dom][0] / kinds[dom][1]
        achieved = alg_bytes[dom] / (avg_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peaks["hbm"],
                    "unit": "GB/s", "frac": achieved / peaks["hbm"], "traffic": traffic,
                    "avg_launch_ms": avg_ms, "launches": kinds[dom][1], "peak_source": peaks["source"],
                    "share_of_step": kinds[dom][0] / ms_eager,
                    "note": "kernel timed with cudaEvent pairs in an eager pass of the same K steps run "
                            "right after the (CUDA-graph) timed region"}
    step_tflops = f_step(H * W) * value / 1e12
    tpeak = (peaks["bf16_sustained"] or peaks["bf16"]) * world
    roofline_step = {"bound": "tensor", "achieved": step_tflops, "peak": tpeak,
                     "unit": "TFLOP/s", "frac": step_tflops / tpeak,
                     "frac_of_burst_peak": step_tflops / (peaks["bf16"] * world),
                     "note": "whole fit step, algorithmic FLOPs (SURVEY.md §8d) vs measured cuBLAS bf16 "
                             "(sustained) — fp16 tcgen05 MMA has the same peak"}
    kernel_ms = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
                 for k, v in kinds.items()}

    cpu_val, cpu_desc, cores = (None, None, None)
    if world == 1 and not args.no_cpu_baseline:
        cpu_val, cpu_desc, cores = cpu_port_steps_per_sec(budget_s=20.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate",
        "data": "synthetic", "config": workload_config(world),
        "exchange": None if world == 1 else ("library peer-memory kernel (NVLink P2P, CUDA IPC)"
                                             if fitter.peer_exchange else "NCCL all-reduce"),
        "final_loss": float(losses[-1].item()), "cuda_graph": bool(graph_mode and fitter._graph is not None),
        "ms_per_step_eager_profiled": ms_eager / args.steps,
        "roofline": roofline, "roofline_step": roofline_step, "kernel_ms": kernel_ms,
        "cpu_baseline": None if cpu_val is None else
        {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_desc},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global DEPTH, HIDDEN, H, W
    DEPTH, HIDDEN, H, W = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
