#!/bin/bash
# multi-GPU bench lines on one box (torchrun, one rank per GPU).  usage (under gpurun --gpus G): bash tools/gpu_scale.sh <tag> <G>
set -u
TAG=${1:-scale}
G=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $G"
timeout 900 $RUN --steps 20 --warmup 3 --no-latency --no-cpu-baseline > $OUT/bench_${TAG}_${G}gpu.json 2> $OUT/bench_${TAG}_${G}gpu.err; echo "N20 rc=$?"
timeout 900 $RUN --steps 8 --warmup 2 --horizon 100 --batch 2048 --no-latency --no-cpu-baseline > $OUT/bench_${TAG}_${G}gpu_N100.json 2> $OUT/bench_${TAG}_${G}gpu_N100.err; echo "N100 rc=$?"
timeout 600 python -m pytest tests -m gpu -q -k "shard or other_device or distributed" > $OUT/pytest_${TAG}_${G}gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_${TAG}_${G}gpu.log
for f in $OUT/bench_${TAG}_${G}gpu.json $OUT/bench_${TAG}_${G}gpu_N100.json; do python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print("$f", d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d["converged_frac"], d["config"]["global_batch"], d["ms_per_step"])
except Exception as e: print("$f ERR", e)
PY
done
