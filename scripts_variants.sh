for f in variants/*.so; do
DD_LIB=$PWD/$f python bench.py --steps 20 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$f', round(d['ms_per_step'],3), 'predict', round(k['k_predict']['ms_per_step'],3), 'cl', round(k['k_assemble<cl>']['ms_per_step'],3), 'cd', round(k['k_assemble<cd>']['ms_per_step'],3))"
done
