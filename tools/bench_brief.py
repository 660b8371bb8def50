"""One-screen digest of a bench.py JSON line: python tools/bench_brief.py FILE [FILE ...]"""
import json
import sys

for path in sys.argv[1:]:
    with open(path) as f:
        d = json.loads(f.read().strip().splitlines()[-1])
    sol = d["config"].get("solver", {})
    print(path, d["value"], d["ms_per_step"], sol.get("sweeps"), sol.get("passes"), sol.get("kernels"))
    print({k: round(v["ms_per_step"], 3) for k, v in d["roofline"].get("kernels", {}).items()})
    print(d.get("check"))
