#!/bin/bash
# A/B sweeps of launch-plan switches on one box: tools/ab_bench.sh <out-prefix> "ENV1=a ENV2=b" "ENV1=c" ...
# each argument is one environment; prints value / ms_per_step / kernel_ms of bench.py --no-cpu-baseline
out=$1; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --quick --steps ${STEPS:-50} --warmup 5 $BENCH_ARGS > gpurun_out/${out}_$i.json 2> gpurun_out/${out}_$i.err
  python - "$envs" gpurun_out/${out}_$i.json <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[2]).read().strip().split("\n")[-1])
    km = {k: round(v["ms_per_step"] * 1000, 1) for k, v in r["kernel_ms"].items()}
    print(f"{sys.argv[1]:45s} {r['value']:8.1f} steps/s  {r['ms_per_step']*1000:7.1f} us  e2e {r['e2e']['value']:.0f}  loss {r['final_loss']:.6f}  {km}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
