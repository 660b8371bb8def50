N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['solver']['sweeps'], d['config']['solver']['retries'], d['e2e'])
print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
PY
