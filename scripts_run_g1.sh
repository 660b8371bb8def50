python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err; tail -3 gpurun_out/bench_g1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_g1.json')); print(d['value'], d['ms_per_step'], d['config']['solver']); print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
DD_NO_GUESS=1 python bench.py --steps 20 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('noguess', d['value'], d['ms_per_step'], d['config']['solver']['sweeps'])"
