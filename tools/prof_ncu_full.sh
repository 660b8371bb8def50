python bench.py --steps 10 --warmup 5 > gpurun_out/r01f_bench.json 2> gpurun_out/r01f_bench.err || exit 1
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:k_eval_sources|k_assemble_c._march|k_correct|k_predict_march|k_rbsor_reg' --launch-skip 10 -c 10 -f -o gpurun_out/r01f_step python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
