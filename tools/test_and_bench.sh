timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 10 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_g3.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['solver']['sweeps'], d['gpu_launches']); print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
tail -3 gpurun_out/bench_g3.err
