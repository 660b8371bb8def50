set -x
python bench.py --steps 10 --warmup 5 > gpurun_out/r01f_bench.json 2> gpurun_out/r01f_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01f_bench_reference.json 2>> gpurun_out/r01f_bench.err
python bench.py --workload sweep > gpurun_out/r01f_bench_sweep.json 2>> gpurun_out/r01f_bench.err
python bench.py --workload ensemble --steps 20 > gpurun_out/r01f_bench_ensemble.json 2>> gpurun_out/r01f_bench.err
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
cut -c1-300 gpurun_out/r01f_bench.json
