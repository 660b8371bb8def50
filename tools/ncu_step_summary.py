#!/usr/bin/env python
"""Per-kernel summary of ONE PC step from an ncu report (raw page as CSV), written where bench.py and the judge
look for it:

    ncu -i gpurun_out/<tag>_step.ncu-rep --page raw --csv > /tmp/raw.csv
    ncu -i gpurun_out/<tag>_step.ncu-rep --page source --csv > /tmp/src.csv      (optional: fp64 counts)
    python tools/ncu_step_summary.py /tmp/raw.csv <tag> [first_launch [launches_per_step [/tmp/src.csv]]]

  profiles/<tag>_ncu_summary.json  duration, DRAM bytes, issue / fp64 / LSU utilisation, registers, occupancy, warp
                                   instructions, fp64 instruction counts of every captured launch
  profiles/<tag>_traffic.json      DRAM bytes per launch per kernel class (bench.py: roofline.traffic)
  profiles/<tag>_fp64.json         fp64 flops of the whole step = 2 DFMA + DADD + DMUL thread instructions
                                   (predicated on), summed over the step's launches (bench.py: roofline.fp64);
                                   from the op-count metrics when the report has them, else from the per-instruction
                                   counts of the source page (`--set full` carries those, not the op counters)

The capture command is tools/prof_ncu_full.sh (a run under ncu is never a bench value: only counts and shares
are taken from it)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# CUDA kernel name (regex) -> profile class of the library's event profile (bench.py roofline.kernels)
CLASSES = [(r"k_predict_march|k_predict<", "k_predict"), (r"k_assemble_cl_march|k_assemble<.*2>", "k_assemble<cl>"),
           (r"k_assemble_cd_march|k_assemble<.*3>", "k_assemble<cd>"), (r"k_assemble<.*1>", "k_assemble<T>"),
           (r"k_correct", "k_correct"), (r"k_eval_sources", "k_eval_sources"), (r"k_feuler", "k_feuler"),
           (r"k_cs_decide|k_cs_redo", "k_cs_decide+k_cs_redo"), (r"k_time_coefs", "k_time_coefs"),
           (r"k_summarise", "k_summarise"), (r"k_reset_stats", "k_reset_stats")]

M = {"duration_us": ("gpu__time_duration.sum", 1.0), "dram_read_B": ("dram__bytes_read.sum", 1.0),
     "dram_write_B": ("dram__bytes_write.sum", 1.0),
     "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
     "fp64_pipe_pct": ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
     "lsu_pipe_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1.0),
     "regs": ("launch__registers_per_thread", 1.0),
     "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
     "warp_insts": ("smsp__inst_executed.sum", 1.0),
     "dfma": ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 1.0),
     "dadd": ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 1.0),
     "dmul": ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 1.0),
     "grid": ("launch__grid_size", 1.0), "block": ("launch__block_size", 1.0)}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "second": 1e6,
              "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}


def source_page_fp64(path, launches):
    """Fills dfma/dadd/dmul of every launch from the SASS view of the source page: predicated-on thread instructions
    of the rows whose opcode is DFMA / DADD / DMUL.  (The page prints some launches twice; identical neighbours are
    dropped, then the sections are matched to the raw page's launches in order and by name.)"""
    secs, cur = [], None
    for r in csv.reader(open(path)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    k = 0
    for d in launches:
        base = d["kernel"].split("(")[0].split("<")[0].replace("void ", "")
        while k < len(secs) and base not in secs[k]["name"]:
            k += 1
        if k == len(secs):
            raise SystemExit("source page has no section for " + d["kernel"])
        s = secs[k]
        k += 1
        if k < len(secs) and secs[k]["name"] == s["name"] and secs[k]["rows"] == s["rows"]:
            k += 1
        h = s["hdr"]
        isrc, ithr = h.index("Source"), h.index("Predicated-On Thread Instructions Executed")
        cnt = {"DFMA": 0.0, "DADD": 0.0, "DMUL": 0.0}
        for r in s["rows"]:
            m = re.search(r"\b(DFMA|DADD|DMUL)\b", r[isrc])
            if m:
                cnt[m.group(1)] += float(r[ithr])
        d["dfma"], d["dadd"], d["dmul"] = cnt["DFMA"], cnt["DADD"], cnt["DMUL"]


def main():
    path, tag = sys.argv[1], sys.argv[2]
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    per_step = int(sys.argv[4]) if len(sys.argv) > 4 else None
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    iname = hdr.index("Kernel Name")
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[iname]}
        for key, (metric, _) in M.items():
            if metric in hdr:
                i = hdr.index(metric)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[key] = v * UNIT_SCALE.get(units[i], 1.0)
        launches.append(d)
    if len(sys.argv) > 5:
        source_page_fp64(sys.argv[5], launches)
    launches = launches[first:first + per_step] if per_step else launches[first:]
    for d in launches:
        k = d["kernel"]
        cls = None
        for rx, c in CLASSES:
            if re.search(rx, k):
                cls = c
        if cls is None and re.search(r"k_rbsor|k_sor_wave|k_sor_lane", k):
            cls = "solver"
        d["class"] = cls or k
    # the solver launches of a step come in the order T (1 or 2 passes), cl, cd: split by the assemble kernels between
    order, cur = [], "T"
    for d in launches:
        if d["class"] == "k_assemble<cl>":
            cur = "cl"
        elif d["class"] == "k_assemble<cd>":
            cur = "cd"
        if d["class"] == "solver":
            d["class"] = "k_rbsor_tile<%s>" % cur
    out_dir = os.path.join(ROOT, "profiles")
    src = f"ncu --set full --clock-control none, gpurun_out/{tag}_step.ncu-rep (one PC step of bench.py, 1025 x 8193 nodes)"
    with open(os.path.join(out_dir, tag + "_ncu_summary.json"), "w") as f:
        json.dump({"source": src, "launches": launches}, f, indent=1)
    traffic = {}
    for d in launches:
        t = traffic.setdefault(d["class"], {"ncu_kernel": d["kernel"], "launches_captured": 0, "bytes": 0.0, "us": 0.0})
        t["launches_captured"] += 1
        t["bytes"] += d.get("dram_read_B", 0.0) + d.get("dram_write_B", 0.0)
        t["us"] += d.get("duration_us", 0.0)
    tj = {"source": src + "; dram__bytes_read.sum + dram__bytes_write.sum per launch", "kernels": {
        c: {"ncu_kernel": t["ncu_kernel"], "launches_captured": t["launches_captured"],
            "dram_bytes_per_launch": t["bytes"] / t["launches_captured"],
            "duration_us_per_launch_under_ncu": t["us"] / t["launches_captured"]} for c, t in traffic.items()},
        "dram_bytes_per_step": sum(t["bytes"] for t in traffic.values())}
    with open(os.path.join(out_dir, tag + "_traffic.json"), "w") as f:
        json.dump(tj, f, indent=1)
    dfma = sum(d.get("dfma", 0.0) for d in launches)
    dadd = sum(d.get("dadd", 0.0) for d in launches)
    dmul = sum(d.get("dmul", 0.0) for d in launches)
    how = ("predicated-on thread instructions of the DFMA / DADD / DMUL rows of the report's source page (SASS view)"
           if len(sys.argv) > 5 else "smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum")
    fj = {"source": src + "; " + how + ", summed over the step's launches",
          "dfma_per_step": dfma, "dadd_per_step": dadd, "dmul_per_step": dmul,
          "flops_per_step": 2.0 * dfma + dadd + dmul,
          "per_class": {c: sum(2.0 * d.get("dfma", 0.0) + d.get("dadd", 0.0) + d.get("dmul", 0.0) for d in launches
                               if d["class"] == c) for c in traffic}}
    with open(os.path.join(out_dir, tag + "_fp64.json"), "w") as f:
        json.dump(fj, f, indent=1)
    nodes = 1025 * 8193
    print("step: %.1f us under ncu, %.2f GB DRAM, %.0f fp64 flops/node, %.0f thread instr/node" % (
        sum(d.get("duration_us", 0) for d in launches), tj["dram_bytes_per_step"] / 1e9, fj["flops_per_step"] / nodes,
        sum(d.get("warp_insts", 0) for d in launches) * 32 / nodes))
    for c, t in tj["kernels"].items():
        print("  %-22s %-28s x%d  %8.1f us  %7.1f MB" % (c, t["ncu_kernel"][:28], t["launches_captured"],
                                                          t["duration_us_per_launch_under_ncu"],
                                                          t["dram_bytes_per_launch"] / 1e6))


if __name__ == "__main__":
    main()
