#!/usr/bin/env python
"""Per-step summary of an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv):
mean duration, share of the step and DRAM bytes of every pipeline kernel.  tools/launch_summary.py launches.csv > summary.json"""
import csv, json, re, sys
from collections import defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) >= 15 and r[0].isdigit()]
per = defaultdict(lambda: defaultdict(float))
count = defaultdict(set)
for r in rows:
    name, metric, val = r[4], r[12], float(r[14].replace(",", ""))
    m = re.search(r"(bk_\w+(<[^>]*>)?|stree_search_\w+(<[^>]*>)?|sa_search_\w+(<[^>]*>)?)", name)
    if not m:
        continue
    k = m.group(1)
    count[k].add(r[0])
    unit = r[13]
    if metric == "gpu__time_duration.sum":
        val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    elif unit in ("Kbyte", "Mbyte", "Gbyte", "byte"):
        val *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    per[k][metric] += val
# the pipeline's kernels: round 1 = rank/colsum/plan/offsets/move/search, round 2 = part/items/search2/unperm
bk = {k: v for k, v in per.items() if re.match(r"bk_(rank|colsum|plan|offsets|move|search|part|items|unperm)", k) and "unsigned long" not in k}  # (the index-output pass belongs to verification calls)
steps = max([len(count[k]) for k in bk if k.startswith("bk_search")] or [1])  # one search launch per pipeline run
out = {"per_kernel": {}, "pipeline_runs_in_capture": steps}
step_us = sum(v["gpu__time_duration.sum"] for v in bk.values()) / steps
for k, v in per.items():
    n = len(count[k])
    out["per_kernel"][k] = {"launches": n, "launches_per_step": n / steps, "mean_us": v["gpu__time_duration.sum"] / n,
                            "share_of_step": (v["gpu__time_duration.sum"] / steps) / step_us if k in bk else None,
                            "dram_read_bytes_per_launch": v["dram__bytes_read.sum"] / n, "dram_write_bytes_per_launch": v["dram__bytes_write.sum"] / n}
out["pipeline_us_per_step"] = step_us
out["pipeline_dram_bytes_per_step"] = sum(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for v in bk.values()) / steps
print(json.dumps(out, indent=1))
