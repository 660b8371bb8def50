python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py --workload ensemble --members 12500 --steps 20 --warmup 3 > gpurun_out/bench_ens.json 2> gpurun_out/bench_ens.err; tail -3 gpurun_out/bench_ens.err; cut -c1-600 gpurun_out/bench_ens.json
python bench.py --workload sweep > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err; tail -3 gpurun_out/bench_sweep.err; cut -c1-900 gpurun_out/bench_sweep.json
