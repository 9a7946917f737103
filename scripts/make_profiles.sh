#!/bin/bash
# Turn the ncu reports of scripts/gpu_profile.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/.
# Usage: bash scripts/make_profiles.sh <tag> <round-prefix>     e.g.  bash scripts/make_profiles.sh r01k r01
TAG=$1; R=${2:-r01}
G=gpurun_out; P=profiles
mkdir -p $P /tmp/cub
(cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all $OLDPWD/stereovision-slam_b200/libsvslam.so > /dev/null)
python scripts/launch_summary.py $G/${TAG}_launches.csv \
  "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python bench.py --profile-window --warmup 3 --streams 256 --groups 1 --steps 45" \
  > $P/${R}_launches_steady_256streams.csv
{
  echo "# ncu --set full --clock-control none --import-source on, one steady-state step of 256 streams (per-frame kernels) and the"
  echo "# keyframe-only kernels of a 45-step window; mean over the captured launches.  dram_*_MB are per launch."
  python scripts/ncu_summary.py $G/${TAG}_full_frame.ncu-rep
  python scripts/ncu_summary.py $G/${TAG}_full_kf.ncu-rep | tail -n +2
} > $P/${R}_ncu_full_summary.csv
python scripts/ncu_lines.py $G/${TAG}_full_kf.ncu-rep k_ba_window /tmp/cub/ba.sm_100a.cubin k_ba_window 30 > $P/${R}_lines_k_ba_window.txt
python scripts/ncu_lines.py $G/${TAG}_full_frame.ncu-rep k_lk_track /tmp/cub/lk.sm_100a.cubin k_lk_trackILi11 30 > $P/${R}_lines_k_lk_track.txt
python scripts/ncu_lines.py $G/${TAG}_full_frame.ncu-rep k_pose_only_lm /tmp/cub/geom.sm_100a.cubin k_pose_only_lm 30 > $P/${R}_lines_k_pose_only_lm.txt
ls -la $P
