#!/bin/bash
# Multi-GPU check (one gpurun --gpus N call): the headline bench under torchrun + the sharded BA solver across N GPUs.
# Usage: gpurun --gpus N --timeout 1200 -- 'bash scripts/gpu_multi.sh <tag> N'
TAG=${1:-r02}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
nproc > $OUT/${TAG}_multi${N}_gpu.txt; nvidia-smi -L >> $OUT/${TAG}_multi${N}_gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/ba_shard_multi.py > $OUT/${TAG}_ba_shard_${N}gpu.json 2> $OUT/${TAG}_ba_shard_${N}gpu.log
echo "ba_shard exit $?"; cat $OUT/${TAG}_ba_shard_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 $BENCH_ARGS > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.log
echo "bench exit $?"
if [ -n "$SECOND_ARGS" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 20 --warmup 5 $SECOND_ARGS > $OUT/${TAG}_bench_${N}gpu_b.json 2> $OUT/${TAG}_bench_${N}gpu_b.log
echo "bench (second: $SECOND_ARGS) exit $?"
fi
python - <<PY
import json, os
for f in ("$OUT/${TAG}_bench_${N}gpu.json", "$OUT/${TAG}_bench_${N}gpu_b.json"):
    if not os.path.exists(f):
        continue
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print("N=$N value %.0f e2e %.0f ms/step %.2f groups %s host_cores %s/%s wait %s cpu_ms/step %s lost %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["context_groups"], d["detail"]["host_cores_this_rank"], d["detail"]["host_cores"], d["detail"].get("host_wait_mode"), d["detail"].get("host_cpu_ms_per_step"), d["detail"]["lost_streams"]))
        print(" phases", d["detail"]["phase_seconds"])
        print(" ba4", d["detail"]["ba_config4"])
    except Exception as e:
        print("parse failed", f, e)
PY
tail -5 $OUT/${TAG}_bench_${N}gpu.log
