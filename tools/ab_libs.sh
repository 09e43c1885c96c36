#!/bin/bash
# A/B of library BUILD variants on one box: tools/ab_libs.sh <out-prefix> "<lib.so or ->|ENV=a ENV=b" ...
# ('-' = the in-tree library); each run swaps the variant in, runs bench.py --quick and prints one line.
out=$1; shift
main=implicit_image_compression_b200/libsirenb200.so
cp $main build/lib_main_backup.so
i=0
for spec in "$@"; do
  i=$((i+1))
  lib=${spec%%|*}; envs=${spec#*|}
  if [ "$lib" = "-" ]; then cp build/lib_main_backup.so $main; else cp $lib $main; fi
  env $envs python bench.py --quick --steps ${STEPS:-50} --warmup 5 $BENCH_ARGS > gpurun_out/${out}_$i.json 2> gpurun_out/${out}_$i.err
  python - "$spec" gpurun_out/${out}_$i.json <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[2]).read().strip().split("\n")[-1])
    km = {k: round(v["ms_per_step"] * 1000, 1) for k, v in r["kernel_ms"].items()}
    print(f"{sys.argv[1]:55s} {r['value']:8.1f} steps/s  {r['ms_per_step']*1000:7.1f} us  e2e {r['e2e']['value']:.0f}  loss {r['final_loss']:.6f}  {km}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
cp build/lib_main_backup.so $main
