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
        sizes = [m << 20 for m in (16, 32, 48, 56, 64, 72, 80, 96, 112, 128, 160, 192, 256, 384, 512, 768, 1024, 2048, 4096, 8192, 16384)]
        for bytes_ in sizes:
            for lanes in (2,):
                gbs = L.sst_probe_gather64(0, bytes_, 200_000_000, lanes, 3)
                emit(kind="probe", bytes=bytes_, lanes=lanes, gbs=gbs, gnodes_per_s=gbs / 64)
    if what in ("all", "tree"):
        g = torch.Generator(device=dev).manual_seed(1)
        n = 1 << logn
        keys = torch.randint(0, MAX, (n,), dtype=torch.int32, device=dev, generator=g)
        keys[0] = MAX
        keys = torch.sort(keys).values.contiguous()
        os.environ['SST_PERSIST'] = '100'  # configure the persisting-L2 carve-out at build time
        tree = sst.STree16.new_params(keys, True, False, False)
        qs = torch.randint(0, MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
        out = torch.empty_like(qs)
        configs = []
        for scheme, G, T in ((3, 2, 2), (5, 2, 1), (5, 2, 2), (5, 4, 1), (5, 4, 2)):
            for hints in (3, 0, 1, 2):
                for persist in (0, 100, 60):
                    configs.append((scheme, G, T, hints, persist))
        for scheme, G, T, hints, persist in configs:
            os.environ.update(SST_T=str(T), SST_TABLE_G=str(G), SST_HINTS=str(hints), SST_PERSIST=str(persist))
            ms = L.sst_time_query_device(tree._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, scheme, 2, 5)
            emit(kind="tree", logn=logn, nq=nq, scheme=scheme, G=G, T=T, hints=hints, persist=persist, ms=ms,
                 gqps=nq / ms / 1e6 if ms > 0 else None, err=L.sst_last_error().decode() if ms < 0 else "")


def ncu_mode():
    """Launch a few configurations once each (run under `ncu -k regex:stree_search_fast`)."""
    L = sst.lib()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    n, nq = 1 << 28, 100_000_000
    keys = torch.randint(0, MAX, (n,), dtype=torch.int32, device=dev, generator=g)
    keys[0] = MAX
    keys = torch.sort(keys).values.contiguous()
    os.environ["SST_PERSIST"] = "100"
    tree = sst.STree16.new_params(keys, True, False, False)
    qs = torch.randint(0, MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
    out = torch.empty_like(qs)
    for scheme, hints, persist in ((5, 3, 0), (5, 7, 0), (5, 7, 100), (5, 6, 100), (5, 7, 200), (5, 4, 100)):
        os.environ.update(SST_T="2", SST_TABLE_G="2", SST_HINTS=str(hints), SST_PERSIST=str(persist))
        for _ in range(2):
            ms = L.sst_time_query_device(tree._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, scheme, 0, 1)
        emit(kind="ncu_cfg", scheme=scheme, hints=hints, persist=persist, ms=ms)


def sizes_mode():
    """Config C2: size sweep 2^10..2^30 keys, 10^8 queries, plain left-max tree (AUTO kernel) and the
    partitioned layouts through the generic kernel."""
    L = sst.lib()
    dev = torch.device("cuda", 0)
    nq = int(os.environ.get("SWEEP_NQ", "100000000"))
    g = torch.Generator(device=dev).manual_seed(3)
    qs = torch.randint(0, MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
    out = torch.empty_like(qs)
    lo, hi = int(os.environ.get("SWEEP_LO", "10")), int(os.environ.get("SWEEP_HI", "30"))
    for logn in range(lo, hi + 1, 2):
        n = 1 << logn
        keys = torch.randint(0, MAX, (n,), dtype=torch.int32, device=dev, generator=g)
        keys[0] = MAX
        keys = torch.sort(keys).values.contiguous()
        tree = sst.STree16.new_params(keys, True, False, False)
        for scheme, name in ((0, "auto"), (5, "table"), (3, "group2"), (4, "generic")):
            ms = L.sst_time_query_device(tree._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, scheme, 2, 5)
            emit(kind="size", logn=logn, layout="stree16_left_max", layers=tree.layers(), kernel=name, ms=ms, gqps=nq / ms / 1e6 if ms > 0 else None)
        del tree
        if logn >= 20:
            for cls, name in ((sst.PartitionedSTree16M, "map"), (sst.PartitionedSTree16C, "compact"), (sst.PartitionedSTree16L, "l1"),
                              (sst.PartitionedSTree16O, "overlap"), (sst.PartitionedSTree16, "simple")):
                t = cls.try_new(keys, 20)
                if t is None:
                    emit(kind="size", logn=logn, layout=name, none=True)
                    continue
                ms = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 0, 1, 3)
                emit(kind="size", logn=logn, layout=name, layers=t.layers(), size_mb=t.size() / 2**20, params=t.params, kernel="auto",
                     ms=ms, gqps=nq / ms / 1e6 if ms > 0 else None)
                del t
        del keys


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sizes":
        sizes_mode()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ncu":
        ncu_mode()
        sys.exit(0)
    main()
