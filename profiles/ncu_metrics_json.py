#!/usr/bin/env python
"""ncu -i <rep> --page raw --csv  ->  the per-launch metric summary bench.py reads (profiles/ncu_r01_*_full_metrics.json).
Usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python profiles/ncu_metrics_json.py "<source note>" > profiles/ncu_....json"""
import csv
import json
import re
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
out = {"source": sys.argv[1] if len(sys.argv) > 1 else "", "launches": []}
for r in rows[2:]:
    name = re.sub(r"^void\s+", "", r[col["Kernel Name"]])
    name = re.sub(r"[<(].*", "", name).strip()
    l = {"kernel": name}
    for k in KEEP:
        if k not in col:
            continue
        v, u = float(r[col[k]].replace(",", "")), units[col[k]]
        if k.startswith("dram__bytes"):
            v *= SCALE.get(u, 1.0); u = "byte"
        if k == "gpu__time_duration.sum":
            v *= SCALE.get(u, 1.0); u = "us"
        l[k] = v
        if u:
            l[k + ".unit"] = u
    l["dram_bytes_total"] = l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    out["launches"].append(l)
json.dump(out, sys.stdout, indent=1)
