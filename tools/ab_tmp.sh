set -u
OUT=gpurun_out
for rep in 1 2; do
for L in fault-tolerant-mpc_b200/csrc/libftmpc.so build/exp/lib_rollchol.so; do
  FTMPC_LIB=$PWD/$L timeout 600 python bench.py --steps 6 --warmup 3 --no-latency --no-cpu-baseline > $OUT/tmp.json 2>$OUT/tmp.err
  python -c "import json; d=json.load(open('$OUT/tmp.json')); r=d['roofline']; print('$L', round(d['value']), round(d['e2e']['value']), r['kernel_ms']['k_solve'], 'chol', r['phase_share']['cholesky'], 'cond', r['phase_share']['cond_blk'], d['converged_frac'])"
done
done
FTMPC_LIB=$PWD/build/exp/lib_rollchol.so timeout 900 python -m pytest tests -m gpu -x -q -k "golden or kkt" > $OUT/pytest_r02l.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_r02l.log
