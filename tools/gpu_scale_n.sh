G=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $G --steps 10 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/bench_r03s_${G}gpu.json 2> gpurun_out/bench_r03s_${G}gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r03s_${G}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['ms_per_step'])"
