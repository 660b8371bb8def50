# usage (on the GPU box): bash tools/prof_bench_and_launch_list.sh <tag>
# The bench lines of one round (default run, reference arm, side workloads) and the ncu launch list of the default
# command; everything lands in gpurun_out/<tag>_*, to be copied into profiles/ by hand.
TAG="${1:-run}"
set -x
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --workload sweep > gpurun_out/${TAG}_bench_sweep.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --workload ensemble --steps 20 > gpurun_out/${TAG}_bench_ensemble.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --integrator feuler --steps 20 > gpurun_out/${TAG}_bench_feuler.json 2>> gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > /dev/null 2>&1 || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > gpurun_out/${TAG}_ncu1.log 2>&1
tail -2 gpurun_out/${TAG}_ncu1.log
cut -c1-300 gpurun_out/${TAG}_bench.json
