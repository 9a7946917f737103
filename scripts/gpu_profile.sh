#!/bin/bash
# ncu evidence of the bench's own configuration (B200_PROFILING.md recipe).  One gpurun call:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh r02'
# (1) launch list (gpu__time_duration.sum) of steady-state steps of the 4096-stream pipeline, one context group
# (2) ncu --set full of the kernels that carry the step (+ the sharded-BA solver, StereoBM)
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
STREAMS=${STREAMS:-4096}
COMMON="bench.py --profile-window --streams $STREAMS --groups 1 --steps 2 --warmup 3 --no-cpu-baseline --configs= --no-ba4 --no-latency --sampler none"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches.csv \
    python $COMMON > $OUT/${TAG}_launches_bench.log 2>&1
echo "launch list exit $?"
if [ -z "$SKIP_FULL" ]; then
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_lk_track|k_ba_window|k_pose_only_lm|k_pyr_down|k_corner_response|k_corner_greedy|k_half_nearest|k_trk_' -c 40 \
    -o $OUT/${TAG}_full python $COMMON > $OUT/${TAG}_full_bench.log 2>&1
echo "full exit $?"
ncu --set full --clock-control none --import-source on -k regex:'k_bs_lm' -c 2 -o $OUT/${TAG}_full_bs \
    python scripts/ba_shard_multi.py > $OUT/${TAG}_full_bs.log 2>&1
echo "full bs exit $?"
fi
