#!/usr/bin/env python
"""Emit GPU measurements in the reference's `Result` JSON schema
(static-search-tree/src/bin/bench.rs:519-545: params, scheme, size, index_size, queries, threads, run,
duration{secs,nanos}, latency [ns/query], layers, cycles, freq) so that the reference's plot.py
(:284-324) can plot GPU rows next to its CPU rows.  Mirrors the CLI of bench.rs:26-46:
  --from/--to (log2 of the input size in BYTES), --queries, --runs, --range, --positive, --dense.
`threads` carries the number of GPUs; `freq` is the SM clock, `cycles` = latency * freq.
Writes results/gpu-results.json (like bench.rs:474-485)."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import torch

import sst_b200 as sst

MAX = sst.MAX


def sizes(lo, hi, dense):
    out = []
    for p in range(lo, hi):
        x = 1 << p
        out += [x, x * 5 // 4, x * 6 // 4, x * 7 // 4] if dense else [x]  # bench.rs:455-472
    return out + [1 << hi]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--from", dest="lo", type=int, default=12)
    ap.add_argument("--to", dest="hi", type=int, default=30)
    ap.add_argument("--queries", type=int, default=10_000_000)
    ap.add_argument("--runs", type=int, default=1)
    ap.add_argument("--range", action="store_true")
    ap.add_argument("--positive", action="store_true")
    ap.add_argument("--dense", action="store_true")
    ap.add_argument("--human", metavar="FASTA", help="keys = 16-mers of this FASTA file (bench.rs:60-76) instead of random keys")
    ap.add_argument("--out", default=os.path.join(ROOT, "results", "gpu-results.json"))
    a = ap.parse_args()
    L = sst.lib()
    dev = torch.device("cuda", 0)
    freq = 1.965e9  # clocks.max.sm of the B200 (MEASURED_PEAKS.json sm_max_mhz)
    results = []
    nq = -(-a.queries // 768) * 768  # next_multiple_of(256 * 3), bench.rs:78
    for run in range(a.runs):
        g = torch.Generator(device=dev).manual_seed(100 + run)
        big = sizes(a.lo, a.hi, a.dense)[-1] // 4
        if a.human:  # bench.rs:60-76: rolling 16-mers of the genome, unsorted here (each size sorts its prefix)
            import numpy as np

            codes = sst.read_fasta_file(a.human)
            keys = sst.kmer_keys(codes, k=16, max_keys=big, sort=False)
            if keys.size < big:
                raise SystemExit(f"{a.human}: only {keys.size} k-mers, need {big} for --to {a.hi}")
            vals_all = torch.from_numpy(keys.view(np.int32)).to(dev)
        else:
            vals_all = torch.randint(0, MAX, (big,), dtype=torch.int32, device=dev, generator=g)  # util.rs:31-42
            vals_all[0] = MAX
        for size in sizes(a.lo, a.hi, a.dense):
            n = size // 4
            vals = torch.sort(vals_all[:n]).values.contiguous()
            if a.positive:  # util.rs:23-28
                qs = vals[torch.randint(0, n, (nq,), device=dev, generator=g)].contiguous()
            else:
                qs = torch.randint(0, MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
            exps = [("single", qs)]
            if a.range:  # bench.rs:84: every query q becomes the pair [q, q+1]
                exps.append(("range", torch.stack([qs, (qs + 1).clamp(max=MAX)], 1).reshape(-1).contiguous()))
            builds = [("STree16 left_max", lambda: sst.STree16.new_params(vals, True, False, False))]
            if not a.range:
                builds.append(("PartitionedSTree16M b=20", lambda: sst.PartitionedSTree16M.try_new(vals, 20)))
            for pname, build in builds:
                idx = build()
                for ename, q in exps:
                    out = torch.empty_like(q)
                    if idx is None:
                        results.append(dict(params=pname, scheme="gpu::" + ename, size=size, index_size=2**64 - 1, queries=q.numel(), threads=1,
                                            run=run, duration=dict(secs=0, nanos=0), latency=0.0, layers=0, cycles=0.0, freq=0.0))
                        continue
                    ms = L.sst_time_query_device(idx._h, C.c_void_p(q.data_ptr()), q.numel(), C.c_void_p(out.data_ptr()), None, 0, 1, 3)
                    secs = ms * 1e-3
                    lat = secs * 1e9 / q.numel()
                    results.append(dict(params=pname, scheme="gpu::" + ename, size=size, index_size=idx.size(), queries=q.numel(), threads=1, run=run,
                                        duration=dict(secs=int(secs), nanos=int((secs % 1) * 1e9)), latency=lat, layers=idx.layers(),
                                        cycles=lat * 1e-9 * freq, freq=freq))
                    print(f"size=2^{size.bit_length()-1:<2} {pname:26s} {ename:7s} {lat*1000:8.2f} ps/query  {q.numel()/secs/1e9:7.2f} Gq/s", flush=True)
                del idx
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(results, open(a.out, "w"))
    print("wrote", a.out, len(results), "rows")


if __name__ == "__main__":
    main()
