#!/bin/bash
# ncu evidence for profiles/: launch list + one `--set full` capture of every kernel of ONE steady-state step
# (bench.py --profile-window brackets the steps with cudaProfilerStart/Stop after priming + warm-up).
# Usage: gpurun --timeout 900 -- 'bash scripts/gpu_profile.sh <tag>'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
SMALL="python bench.py --profile-window --warmup 3 --streams 256 --groups 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv $SMALL --steps 8 > $OUT/${TAG}_launches_bench.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -f -o $OUT/${TAG}_full $SMALL --steps 1 > $OUT/${TAG}_full_bench.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
