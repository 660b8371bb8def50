# usage (on the GPU box): bash tools/test_and_bench.sh <tag>   -- GPU tests, then a short bench with the kernel split
TAG="${1:-run}"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_quick.json 2>gpurun_out/${TAG}_bench_quick.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_quick.json").read())
    print(d["value"], d["ms_per_step"], d["config"]["solver"]["sweeps"], d["config"]["solver"]["passes"], d["gpu_launches"])
    print({k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("bench failed:", e)
PY
tail -3 gpurun_out/${TAG}_bench_quick.err
