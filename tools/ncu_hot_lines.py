#!/usr/bin/env python
"""Hottest SASS lines (warp-stall samples) of the first kernel in an .ncu-rep captured with --import-source on."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = rows[2:]
tot = sum(int(r[isamp]) for r in data); totex = sum(int(r[iex]) for r in data)
print("total samples", tot, "total warp instructions", totex)
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:topn]
for i in sorted(top):
    r = data[i]
    print(f"{i:5d} {int(r[isamp]):6d} {100 * int(r[isamp]) / tot:5.1f}%  ex={int(r[iex]):9d}  {r[isrc].strip()[:100]}")
