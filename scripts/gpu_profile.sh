#!/bin/bash
# ncu evidence for profiles/: launch list + `--set full` captures of steady-state steps
# (bench.py --profile-window brackets the steps with cudaProfilerStart/Stop after priming + warm-up).
# Keyframes (detect / right-LK / triangulate / BA) come in bursts, so the keyframe-only kernels are captured from a
# 45-step window with a kernel-name filter, the per-frame kernels from a one-step window.
# Usage: gpurun --timeout 1200 -- 'bash scripts/gpu_profile.sh <tag>'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
SMALL="python bench.py --profile-window --warmup 3 --streams 256 --groups 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv $SMALL --steps 45 > $OUT/${TAG}_launches_bench.log 2>&1
echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -f -o $OUT/${TAG}_full_frame $SMALL --steps 1 > $OUT/${TAG}_full_frame_bench.log 2>&1
echo "ncu full (per-frame kernels) exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_ba_window|k_corner|k_triangulate|k_mask' -c 12 \
    -f -o $OUT/${TAG}_full_kf $SMALL --steps 45 > $OUT/${TAG}_full_kf_bench.log 2>&1
echo "ncu full (keyframe kernels) exit $?"
ls -la $OUT
