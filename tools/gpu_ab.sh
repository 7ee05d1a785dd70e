#!/bin/bash
# quick GPU pass for kernel development: smoke, selected GPU tests, short benches of both QP methods
# usage (under gpurun): bash tools/gpu_ab.sh <tag> ["pytest -k expression"]
set -u
TAG=${1:-dev}
KEXPR=${2:-}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_$TAG.log
if [ -n "$KEXPR" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_$TAG.log
fi
for M in 1 0; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-latency --no-cpu-baseline --solver-opts "{\"qp_method\": $M}" > $OUT/bench_${TAG}_m$M.json 2> $OUT/bench_${TAG}_m$M.err
  echo "bench m$M rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_m$M.json"))
    r=d["roofline"]
    print("m$M value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"conv",d["converged_frac"],"sqp",round(d["sqp_iters_mean"],2),"qp",round(d["qp_iters_mean"],1),"k_ms",r["kernel_ms"],"busy",r["sm_busy_frac"])
    print({k:v for k,v in r["phase_share"].items() if v>0.004})
    print(r["phase_counters"])
except Exception as e:
    print("no bench line",e); print(open("$OUT/bench_${TAG}_m$M.err").read()[-1500:])
PY
done
