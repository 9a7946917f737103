#!/bin/bash
# ncu evidence of the bench's own configuration (B200_PROFILING.md recipe).  One gpurun call:
#   gpurun --timeout 2400 -- 'bash scripts/gpu_profile.sh r02'
# (1) launch list (gpu__time_duration.sum) of steady-state steps of the bench's own configuration (4096 streams, 4 context groups)
# (2) ncu --set full of the kernels that carry the step, of the sharded-BA solver and of StereoBM
# (3) sysmem (PCIe) sector counters of the zero-copy ingest kernel in the e2e configuration
# The .ncu-rep files are summarised ON THE BOX (gpurun brings back at most 64 MiB) and deleted.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT /tmp/cub
STREAMS=${STREAMS:-4096}
GROUPS_=${GROUPS_:-16}
COMMON="bench.py --profile-window --streams $STREAMS --groups $GROUPS_ --steps 2 --warmup 3 --no-cpu-baseline --configs= --no-ba4 --no-latency --sampler none"
(cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all $OLDPWD/stereovision-slam_b200/libsvslam.so > /dev/null)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches.csv \
    python $COMMON > $OUT/${TAG}_launches_bench.log 2>&1
echo "launch list exit $?"
python scripts/launch_summary.py $OUT/${TAG}_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python $COMMON" > $OUT/${TAG}_launches_summary.csv
head -20 $OUT/${TAG}_launches_summary.csv
if [ -z "$SKIP_FULL" ]; then
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_lk_track|k_ba_window|k_pose_only_lm|k_pyr_down|k_corner_response|k_corner_greedy|k_corner_select|k_half_nearest|k_trk_|k_ba_build' -c 96 \
    -o /tmp/${TAG}_full python $COMMON > $OUT/${TAG}_full_bench.log 2>&1
echo "full exit $?"
python scripts/ncu_summary.py /tmp/${TAG}_full.ncu-rep > $OUT/${TAG}_full_summary.csv
python scripts/ncu_lines.py /tmp/${TAG}_full.ncu-rep k_lk_track /tmp/cub/lk.sm_100a.cubin k_lk_trackILi11 40 > $OUT/${TAG}_lines_k_lk_track.txt
python scripts/ncu_lines.py /tmp/${TAG}_full.ncu-rep k_ba_window /tmp/cub/ba.sm_100a.cubin k_ba_window 40 > $OUT/${TAG}_lines_k_ba_window.txt
python scripts/ncu_lines.py /tmp/${TAG}_full.ncu-rep k_pose_only_lm /tmp/cub/geom.sm_100a.cubin k_pose_only_lm 30 > $OUT/${TAG}_lines_k_pose_only_lm.txt
python scripts/ncu_lines.py /tmp/${TAG}_full.ncu-rep k_pyr_down_tma /tmp/cub/images.sm_100a.cubin k_pyr_down_tma 25 > $OUT/${TAG}_lines_k_pyr_down_tma.txt
if [ -z "$SKIP_BS" ]; then
ncu --set full --clock-control none --import-source on -k regex:'k_bs_lm' -c 1 -o /tmp/${TAG}_full_bs \
    python scripts/ba_shard_multi.py > $OUT/${TAG}_full_bs.log 2>&1
echo "full bs exit $?"
python scripts/ncu_summary.py /tmp/${TAG}_full_bs.ncu-rep > $OUT/${TAG}_full_bs_summary.csv
python scripts/ncu_lines.py /tmp/${TAG}_full_bs.ncu-rep k_bs_lm /tmp/cub/ba_shard.sm_100a.cubin k_bs_lm 40 > $OUT/${TAG}_lines_k_bs_lm.txt
fi
if [ -z "$SKIP_BM" ]; then
ncu --set full --clock-control none -k regex:'k_bm_' -c 3 -o /tmp/${TAG}_full_bm python scripts/bm_run.py > $OUT/${TAG}_full_bm.log 2>&1
echo "full bm exit $?"
python scripts/ncu_summary.py /tmp/${TAG}_full_bm.ncu-rep > $OUT/${TAG}_full_bm_summary.csv
fi
ncu --metrics gpu__time_duration.sum,syslts__t_sectors_aperture_sysmem_op_read.sum,syslts__t_sectors_aperture_sysmem.sum,pcie__read_bytes.sum,pcie__write_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --profile-from-start off -k regex:'k_half_nearest' --csv --log-file $OUT/${TAG}_ingest_sysmem.csv \
    python $COMMON --profile-e2e > $OUT/${TAG}_ingest_bench.log 2>&1
echo "ingest exit $?"
rm -f /tmp/${TAG}_full.ncu-rep /tmp/${TAG}_full_bs.ncu-rep /tmp/${TAG}_full_bm.ncu-rep
fi
ls -la $OUT | grep ${TAG}
