#!/bin/bash
# One gpurun call: GPU parity tests + the default bench line (+ optional extra bench variants via EXTRA).
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh <tag>'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
  tail -5 $OUT/${TAG}_pytest_gpu.log
fi
if [ -z "$SKIP_BENCH" ]; then
  timeout 900 python bench.py $BENCH_ARGS > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log
  echo "bench exit $?"
  tail -c 1500 $OUT/${TAG}_bench.json
fi
