for env in "A=1" "DD_SOLVER_PIPE=0" "DD_NO_MARCH=1"; do
env $env timeout 300 python bench.py --workload ensemble --steps 20 2>gpurun_out/bench_g3.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$env', d['value'], d['ms_per_step'], d['config'].get('solver'))"
done
env DD_SOLVER_PIPE=0 timeout 300 python bench.py --workload sweep 2>>gpurun_out/bench_g3.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sweep nopipe', d['value'], d['ms_per_step'])"
tail -3 gpurun_out/bench_g3.err
