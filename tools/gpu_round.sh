#!/bin/bash
# One GPU-box pass: smoke, GPU parity tests, bench (N=1), launch list and one full ncu capture of the top kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [kernel-regex]
set -u
TAG=${1:-r01}
KRE=${2:-k_solve}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu_$TAG.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks_$TAG.csv &
SMI=$!
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
kill $SMI
[ "${SKIP_NCU:-0}" = "1" ] && exit 0
SMALL="python bench.py --batch 592 --steps 1 --warmup 3 --no-latency --no-cpu-baseline"
timeout 600 $SMALL > $OUT/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
timeout 600 $SMALL > $OUT/plain2_$TAG.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o $OUT/prof_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
