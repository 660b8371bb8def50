#!/usr/bin/env bash
# development aid: solver tile-shape sweep on the mesh workload (prints per-kernel ms per step)
for th in 256 512; do for ti in 16 24 32 48 64; do for sj in 64 128; do
  out=$(DD_THREADS=$th DD_TILE_I=$ti DD_STAGE_J=$sj python bench.py --steps 5 --warmup 8 --no-e2e --no-cpu-baseline 2>/dev/null)
  python - "$th" "$ti" "$sj" "$out" <<'PY'
import json, sys
th, ti, sj, out = sys.argv[1:5]
try:
    d = json.loads(out)
    k = d["roofline"]["kernels"]
    print(th, ti, sj, "ms/step %.3f" % d["ms_per_step"], {n: round(v["ms_per_step"], 3) for n, v in k.items() if "rbsor" in n},
          d["config"]["solver"]["sweeps"], d["config"]["solver"]["passes"])
except Exception as e:
    print(th, ti, sj, "failed", e, out[:200])
PY
done; done; done
