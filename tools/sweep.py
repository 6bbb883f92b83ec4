#!/usr/bin/env python
"""GPU tuning sweep (development tool): times every lower_bound kernel variant and the random
64-byte gather probe, writes JSON lines to gpurun_out/sweep.jsonl."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import torch

import sst_b200 as sst

MAX = sst.MAX
out_path = os.path.join(ROOT, "gpurun_out", "sweep.jsonl")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
outf = open(out_path, "a")


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    outf.write(line + "\n")
    outf.flush()


def main():
    L = sst.lib()
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    logn = int(os.environ.get("SWEEP_LOGN", "28"))
    nq = int(os.environ.get("SWEEP_NQ", "100000000"))
    dev = torch.device("cuda", 0)
    if what in ("all", "probe"):
        for bytes_ in (64 << 20, 1 << 30, 8 << 30):
            for lanes in (16, 8, 4, 2):
                gbs = L.sst_probe_gather64(0, bytes_, 200_000_000, lanes, 3)
                emit(kind="probe", bytes=bytes_, lanes=lanes, gbs=gbs, gnodes_per_s=gbs / 64)
    if what in ("all", "tree"):
        g = torch.Generator(device=dev).manual_seed(1)
        n = 1 << logn
        keys = torch.randint(0, MAX, (n,), dtype=torch.int32, device=dev, generator=g)
        keys[0] = MAX
        keys = torch.sort(keys).values.contiguous()
        tree = sst.STree16.new_params(keys, True, False, False)
        qs = torch.randint(0, MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
        out = torch.empty_like(qs)
        configs = []
        for scheme, T in ((1, 1), (1, 2), (2, 1), (3, 1), (3, 2), (4, 1)):
            for smem in (0, 16, 224):
                for hints in (3, 0):
                    for threads in (1024, 512):
                        if scheme == 4 and (smem or hints != 3 or threads != 1024):
                            continue
                        configs.append((scheme, T, smem, hints, threads))
        for scheme, T, smem, hints, threads in configs:
            os.environ.update(SST_T=str(T), SST_SMEM_KB=str(smem), SST_HINTS=str(hints), SST_THREADS=str(threads))
            ms = L.sst_time_query_device(tree._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, scheme, 2, 5)
            emit(kind="tree", logn=logn, nq=nq, scheme=scheme, T=T, smem_kb=smem, hints=hints, threads=threads, ms=ms,
                 gqps=nq / ms / 1e6 if ms > 0 else None, err=L.sst_last_error().decode() if ms < 0 else "")


if __name__ == "__main__":
    main()
