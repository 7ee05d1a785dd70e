set -u
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "long_horizon" > $OUT/pytest_r02k.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_r02k.log
for M in 8 0; do
  timeout 900 python bench.py --horizon 100 --batch 148 --steps 1 --warmup 3 --streams 1 --no-latency --no-cpu-baseline --solver-opts "{\"qp_method\": $M}" > $OUT/bench_r02k_N100_m$M.json 2>$OUT/tmp.err
  python -c "
import json; d=json.load(open('$OUT/bench_r02k_N100_m$M.json')); r=d['roofline']; print('N100 m$M', round(d['value'],1), 'conv', d['converged_frac'], 'sqp', d['sqp_iters_mean'], 'qp', d['qp_iters_mean'], r['kernel_ms']); print({k:v for k,v in r['phase_share'].items() if v>0.01})"
done
timeout 900 python bench.py --horizon 30 --batch 1184 --steps 2 --warmup 3 --no-latency --no-cpu-baseline > $OUT/bench_r02k_N30.json 2>$OUT/tmp.err
python -c "
import json; d=json.load(open('$OUT/bench_r02k_N30.json')); r=d['roofline']; print('N30', round(d['value'],1), 'conv', d['converged_frac'], r['kernel_ms']); print({k:v for k,v in r['phase_share'].items() if v>0.01})"
for rep in 1 2; do
for L in fault-tolerant-mpc_b200/csrc/libftmpc.so build/exp/lib_rollcond.so; do
  FTMPC_LIB=$PWD/$L timeout 600 python bench.py --steps 6 --warmup 3 --no-latency --no-cpu-baseline > $OUT/tmp.json 2>$OUT/tmp.err
  python -c "import json; d=json.load(open('$OUT/tmp.json')); print('$L', round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms']['k_solve'], d['roofline']['phase_share']['cond_blk'])"
done
done
