#!/usr/bin/env python
"""Hottest SASS instructions (warp-stall samples) of one kernel in an .ncu-rep: tools/ncu_hot.py rep kernel-regex [N]."""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for si, s in enumerate(starts[:1]):
    h = rows[s]
    end = starts[si + 1] - 1 if si + 1 < len(starts) else len(rows)
    body = [r for r in rows[s + 1:end] if len(r) == len(h)]
    ci = {n: h.index(n) for n in h}
    tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
    print(f"total samples {tot}, instructions {len(body)}")
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = {n: sum(int(r[ci[n]] or 0) for r in body) for n in stalls}
    print("stall mix:", {k: round(100 * v / max(1, sum(agg.values())), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    idx = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]] or 0))[:topn]
    for i in sorted(idx):
        r = body[i]
        top = sorted(((int(r[ci[n]] or 0), n) for n in stalls), reverse=True)[:2]
        print(f"{i:5d} {100*int(r[ci['# Samples']] or 0)/max(1,tot):5.1f}%  {r[ci['Source']][:70]:70s} {top}")
