# usage (on the GPU box): bash tools/prof_ncu_full.sh <tag>
# One PC step of the default bench under `ncu --set full` (after the same command has run clean without ncu):
# gpurun_out/<tag>_step.ncu-rep, summarised here with tools/ncu_step_summary.py.
TAG="${1:-run}"
timeout 200 python bench.py --steps 2 --warmup 12 --no-e2e --no-cpu-baseline --no-check > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 185 --launch-count 24 -f -o gpurun_out/${TAG}_step python bench.py --steps 2 --warmup 12 --no-e2e --no-cpu-baseline --no-check > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log
