#!/bin/bash
# round-end measurement pass on one GPU: smoke, all GPU tests, the default bench line (with CPU baseline and latency),
# the reference arm, the other configurations, launch list + ncu full capture.  usage: bash tools/gpu_final.sh <tag>
set -u
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks_$TAG.csv &
SMI=$!
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
kill $SMI
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference_arm.json 2> $OUT/bench_${TAG}_reference_arm.err; echo "reference arm rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --batch 1024 --no-latency --no-cpu-baseline > $OUT/bench_${TAG}_b1k.json 2>/dev/null; echo "b1k rc=$?"
timeout 600 python bench.py --steps 3 --warmup 2 --batch 65536 --no-latency --no-cpu-baseline > $OUT/bench_${TAG}_b64k.json 2>/dev/null; echo "b64k rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --horizon 15 --no-latency --no-cpu-baseline > $OUT/bench_${TAG}_N15.json 2>/dev/null; echo "N15 rc=$?"
timeout 600 python bench.py --steps 4 --warmup 2 --horizon 30 --batch 4096 --no-latency --no-cpu-baseline > $OUT/bench_${TAG}_N30.json 2>/dev/null; echo "N30 rc=$?"
timeout 900 python bench.py --steps 4 --warmup 2 --horizon 100 --batch 2048 --no-latency --no-cpu-baseline > $OUT/bench_${TAG}_N100.json 2>/dev/null; echo "N100 rc=$?"
bash tools/gpu_ncu.sh $TAG
