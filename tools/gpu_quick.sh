#!/bin/bash
# quick GPU pass: smoke, (optional) selected GPU tests, short benches at N = 20 / 100 / 30
# usage (under gpurun): bash tools/gpu_quick.sh <tag> ["pytest -k expression"|all] [extra horizons...]
set -u
TAG=${1:-dev}
KEXPR=${2:-}
[ $# -ge 2 ] && shift 2 || shift $#
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_$TAG.log
if [ "$KEXPR" = "all" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_$TAG.log
elif [ -n "$KEXPR" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_$TAG.log
fi
show() {
python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$1 value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"conv",d["converged_frac"],"sqp",round(d["sqp_iters_mean"],2),"qp",round(d["qp_iters_mean"],1),"k_ms",r["kernel_ms"],"busy",r["sm_busy_frac"],"frac",round(r["frac"],4), "lat", (d.get("latency") or {}).get("p50_ms"))
    print({k:v for k,v in r["phase_share"].items() if v>0.004})
    print(r["phase_counters"])
except Exception as e:
    print("no bench line",e); print(open("$2").read()[-1500:])
PY
}
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
show $OUT/bench_$TAG.json $OUT/bench_$TAG.err
for H in "$@"; do
  B=1184; [ "$H" -ge 60 ] && B=296
  timeout 600 python bench.py --steps 2 --warmup 1 --no-latency --no-cpu-baseline --horizon $H --batch $B > $OUT/bench_${TAG}_N$H.json 2> $OUT/bench_${TAG}_N$H.err; echo "bench N=$H rc=$?"
  show $OUT/bench_${TAG}_N$H.json $OUT/bench_${TAG}_N$H.err
done
