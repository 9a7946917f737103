mkdir -p gpurun_out /tmp/cub
(cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all $OLDPWD/stereovision-slam_b200/libsvslam.so > /dev/null)
python scripts/ba_shard_multi.py 2>/dev/null | tail -1
ncu --set full --clock-control none --import-source on -k regex:'k_bs_lm' -c 1 -o /tmp/bs python scripts/ba_shard_multi.py > /dev/null 2>&1
python scripts/ncu_lines.py /tmp/bs.ncu-rep k_bs_lm /tmp/cub/ba_shard.sm_100a.cubin k_bs_lm 45 > gpurun_out/r02n_lines_k_bs_lm.txt
python scripts/ncu_summary.py /tmp/bs.ncu-rep > gpurun_out/r02n_bs_summary.csv
head -50 gpurun_out/r02n_lines_k_bs_lm.txt
