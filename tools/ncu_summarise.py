"""Turns ncu output brought back in gpurun_out/ into the summaries committed under profiles/.
  python tools/ncu_summarise.py launches gpurun_out/launches.csv  > profiles/rNN_ncu_launch_list_summary.txt
  python tools/ncu_summarise.py full gpurun_out/prof.ncu-rep [...] > profiles/rNN_ncu_full_summary.md
      (also rewrites profiles/ncu_traffic.json: DRAM bytes per launch per bench.py kernel kind)
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name[:78]


def launches(path):
    rows = [l for l in open(path) if not l.startswith("==")]
    tot = defaultdict(lambda: [0.0, 0])
    for r in csv.DictReader(io.StringIO("".join(rows))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit == "ns" else v * (1e3 if unit == "ms" else 1.0)
        k = short(r["Kernel Name"])
        tot[k][0] += us
        tot[k][1] += 1
    total = sum(v[0] for v in tot.values())
    n = sum(v[1] for v in tot.values())
    print(f"total kernel time {total:.1f} us over {n} launches")
    for k, (us, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"  {us:9.1f} us  {100 * us / total:5.1f}%  n={c:4d}  avg {us / c:8.1f} us  {k}")


KIND = [("rowgemm_kernel<256, 256, 0", "fwd_gemm"), ("rowgemm_kernel<256, 256, 1", "dx_gemm"),
        ("bwd_merged_kernel<256", "dx_gemm_merged"), ("step_end_kernel", "reduce_partials"),
        ("colgemm_kernel", "dw_gemm"), ("last_layer_tc_kernel", "last_layer_loss"), ("tail_tc_kernel", "last_layer_loss"),
        ("tc_last_layer_kernel", "last_layer_loss"), ("tc_layer0_grad_kernel", "layer0_grad"),
        ("tc_first_layer_kernel", "first_layer")]


def full(paths):
    agg = OrderedDict()
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rd = list(csv.reader(io.StringIO(out)))
        head = rd[0]

        def col(sub):
            for i, h in enumerate(head):
                if h.endswith(sub):
                    return i
            return None
        idx = {k: col(k) for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                   "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                   "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                   "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                                   "launch__registers_per_thread", "smsp__inst_executed.sum")}
        units = rd[1]
        name_i = head.index("Kernel Name")
        for r in rd[2:]:
            if len(r) <= name_i:
                continue
            k = short(r[name_i])
            a = agg.setdefault(k, defaultdict(float))
            a["n"] += 1
            for key, i in idx.items():
                if i is None or r[i] == "":
                    continue
                v = float(r[i].replace(",", ""))
                u = units[i]
                if key == "gpu__time_duration.sum":
                    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
                if key.startswith("dram__bytes"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                a[key] += v
    print("kernel | launches | time us | dram read MB | dram write MB | dram % of peak | tensor pipe % | sm % | "
          "regs | warp-instr M")
    print("---|---|---|---|---|---|---|---|---|---")
    traffic = {}
    for k, a in agg.items():
        n = a["n"]
        g = lambda key: a[key] / n  # noqa: E731
        print(f"{k} | {int(n)} | {g('gpu__time_duration.sum'):.1f} | {g('dram__bytes_read.sum') / 1e6:.1f} | "
              f"{g('dram__bytes_write.sum') / 1e6:.1f} | "
              f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{g('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{g('launch__registers_per_thread'):.0f} | {g('smsp__inst_executed.sum') / 1e6:.1f}")
        for pat, kind in KIND:
            if pat in k:
                t = traffic.setdefault(kind, [0.0, 0])
                t[0] += a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]
                t[1] += n
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump({k: v[0] / v[1] for k, v in traffic.items()}, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
