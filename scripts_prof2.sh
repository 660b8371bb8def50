set -x
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:k_rbsor_reg|k_predict|k_assemble|k_correct|k_eval_sources' --launch-skip 14 -c 14 -f -o gpurun_out/r01c_full python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out/
