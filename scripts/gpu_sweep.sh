#!/bin/bash
# Group-count sweep of the headline bench (one gpurun call).  Usage: gpurun -- 'bash scripts/gpu_sweep.sh <tag> "1 2 4"'
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nproc > $OUT/${TAG}_gpu.txt
for G in ${2:-1 2 4}; do
  timeout 600 python bench.py --groups $G --steps 30 --warmup 5 --no-cpu-baseline --configs "" --no-ba4 --no-latency $EXTRA \
      > $OUT/${TAG}_sweep_g$G.json 2> $OUT/${TAG}_sweep_g$G.log
  echo "G=$G exit $?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_sweep_g$G.json"))
    print("G=$G value %.0f e2e %.0f ms/step %.2f e2e ms %.2f lost %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["detail"]["lost_streams"]))
    print(" phases", d["detail"]["phase_seconds"])
    print(" shares", d["detail"]["kernel_time_share"])
    print(" roof", d["roofline"]["kernel"], d["roofline"]["frac"])
except Exception as e:
    print("parse failed", e)
PY
done
