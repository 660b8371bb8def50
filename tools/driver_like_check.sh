# what the round-end driver does on a fresh box: GPU tests, smoke, default bench lines
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err; echo "reference rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/final_bench.json"))
need = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"]
print("missing keys:", [k for k in need if k not in d])
print(d["value"], d["ms_per_step"], d["steps"], d["warmup"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
print({k: d["roofline"][k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "traffic")})
print(d["cpu_baseline"])
r = json.load(open("gpurun_out/final_bench_reference.json"))
print("reference:", r["impl"], r["value"], r["cpu_baseline"]["cores"], r["e2e"])
PY
