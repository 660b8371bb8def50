python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:k_predict_march' --launch-skip 2 -c 1 -f -o gpurun_out/r01d_march python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
