import json,sys
d=json.loads(sys.stdin.read()); print(sys.argv[1], d["value"], d["ms_per_step"], d["config"]["solver"]["sweeps"], d["config"]["solver"]["passes"]); print({k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()}); print(d.get("check"))
