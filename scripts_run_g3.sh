python bench.py --steps 10 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_g3.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
python bench.py --steps 10 --warmup 5 --no-cpu-baseline --e2e-single 2>gpurun_out/bench_g3.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
tail -3 gpurun_out/bench_g3.err
