#!/bin/bash
# SASS evidence of the Blackwell-native paths: instruction counts per cubin of the shipped library -> profiles/<round>_sass_grep.txt
# Usage: bash scripts/sass_grep.sh r02
R=${1:-r02}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p /tmp/cub_sass && cd /tmp/cub_sass && rm -f *.cubin && cuobjdump -xelf all $ROOT/stereovision-slam_b200/libsvslam.so > /dev/null
{
echo "# SASS evidence of the Blackwell-native staging / synchronisation paths in stereovision-slam_b200/libsvslam.so (sm_100a):"
echo "# cuobjdump -sass per cubin, instruction counts.  UTMALDG = cp.async.bulk.tensor (TMA tile load), SYNCS = mbarrier ops,"
echo "# UBLKCP = cp.async.bulk, LDGSTS = cp.async, REDUX = redux.sync, ELECT = elect-one issue of the TMA copy,"
echo "# sys_scope = system-scope release / acquire / relaxed loads and stores (the peer exchange windows of k_bs_lm), DFMA = FP64 FMA."
echo "cubin,UTMALDG,SYNCS,UBLKCP,LDGSTS,REDUX,ELECT,sys_scope,DFMA"
for f in /tmp/cub_sass/*.cubin; do
  s=$(cuobjdump -sass $f); n=$(basename $f .sm_100a.cubin)
  echo "$n,$(echo "$s" | grep -c UTMALDG),$(echo "$s" | grep -c 'SYNCS'),$(echo "$s" | grep -c UBLKCP),$(echo "$s" | grep -c LDGSTS),$(echo "$s" | grep -c REDUX),$(echo "$s" | grep -c ' ELECT'),$(echo "$s" | grep -cE '\.SYS'),$(echo "$s" | grep -c DFMA)"
done
echo "# kernels that contain UTMALDG / SYNCS:"
for f in /tmp/cub_sass/*.cubin; do cuobjdump -sass $f | awk '/Function :/{fn=$3} /UTMALDG/{c[fn]++} /SYNCS/{d[fn]++} END{for(k in d) print "#   " k ": UTMALDG " c[k]+0 ", SYNCS " d[k]}'; done
} > $ROOT/profiles/${R}_sass_grep.txt
cat $ROOT/profiles/${R}_sass_grep.txt
