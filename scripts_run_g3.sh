for spp in "" 7 5 4; do
DD_SWEEPS_PER_PASS=$spp timeout 300 python bench.py --steps 20 --warmup 10 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_g3.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('spp=$spp', round(d['ms_per_step'],4), d['config']['solver']['sweeps'], d['config']['solver']['passes'], {n:round(k[n]['ms_per_step'],3) for n in k if 'rbsor' in n or 'cs_' in n})"
done
tail -3 gpurun_out/bench_g3.err
