#!/bin/bash
# One gpurun call: the headline bench (quick form: no CPU baseline, no other configs) under several environment variants.
# Usage: gpurun --timeout 900 -- 'bash scripts/gpu_variants.sh <tag> "VAR=1 VAR2=x" "..." ...'   ("-" = no variables)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
Q="--no-cpu-baseline --configs= --no-ba4 --no-latency --no-kernel-pass $BENCH_ARGS"
i=0
for V in "$@"; do
  i=$((i+1))
  [ "$V" = "-" ] && V=""
  PRE=""; ENVV=""
  for tok in $V; do case "$tok" in TASKSET=*) PRE="taskset -c ${tok#TASKSET=}";; ARGS=*) EXTRA="$(echo ${tok#ARGS=} | tr , ' ')";; *) ENVV="$ENVV $tok";; esac; done
  env $ENVV $PRE timeout 300 python bench.py $Q $EXTRA > $OUT/${TAG}_v$i.json 2> $OUT/${TAG}_v$i.log
  python - "$OUT/${TAG}_v$i.json" "$V" <<'PY'
import json, sys
try:
    d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print("%-50s value %8.0f  e2e %8.0f  ms/step %6.2f  groups %s  wait %s" % (sys.argv[2] or "-", d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["context_groups"], d["detail"].get("host_wait_mode")))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
  EXTRA=""
done
