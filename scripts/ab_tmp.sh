#!/bin/bash
# A/B: LK at 8 vs 6 CTAs per SM, pose-only LM at 4 CTAs per SM (launch bounds)
B="python bench.py --steps 40 --warmup 5 --configs= --no-latency --no-cpu-baseline --no-ba4"
pr() { python - "$1" "$2" <<'PY'
import json,sys
d=[json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")][-1]
k=d["detail"]["kernel_ms"]
print(sys.argv[1], "value %.0f e2e %.0f  lk %.1f pose %.1f ba %.1f (ms summed over %d steps)" % (d["value"], d["e2e"]["value"], k["k_lk_track"][0], k["k_pose_only_lm"][0], k["k_ba_window"][0], d["steps"]))
PY
}
$B > gpurun_out/ab_lk8.json 2>/dev/null; pr lk8 gpurun_out/ab_lk8.json
SVS_LK_6=1 $B > gpurun_out/ab_lk6.json 2>/dev/null; pr lk6 gpurun_out/ab_lk6.json
$B > gpurun_out/ab_lk8b.json 2>/dev/null; pr lk8_again gpurun_out/ab_lk8b.json
$B --groups 3 > gpurun_out/ab_g3.json 2>/dev/null; pr groups3 gpurun_out/ab_g3.json
$B --groups 4 > gpurun_out/ab_g4.json 2>/dev/null; pr groups4 gpurun_out/ab_g4.json
sed -i 's/__global__ void __launch_bounds__(PO_WARPS \* 32)/__global__ void __launch_bounds__(PO_WARPS * 32, 4)/' stereovision-slam_b200/csrc/geom.cu
python -c "
import importlib.util,sys
spec=importlib.util.spec_from_file_location('b','stereovision-slam_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build(verbose=False)" > /dev/null 2>&1
$B > gpurun_out/ab_pose4.json 2>/dev/null; pr pose4 gpurun_out/ab_pose4.json
