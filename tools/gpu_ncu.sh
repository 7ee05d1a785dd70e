#!/bin/bash
# one ncu --set full capture of the 4th k_solve launch of a small bench run (592 instances), plus the launch list
# usage (under gpurun): bash tools/gpu_ncu.sh <tag> [kernel-regex] [extra bench args]
set -u
TAG=${1:-dev}
KRE=${2:-k_solve}
[ $# -ge 2 ] && shift 2 || shift $#
OUT=gpurun_out
mkdir -p $OUT
SMALL="python bench.py --batch 592 --steps 1 --warmup 3 --no-latency --no-cpu-baseline $*"
timeout 600 $SMALL > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o $OUT/prof_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT/prof_$TAG.ncu-rep
