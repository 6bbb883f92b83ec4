#!/usr/bin/env python
"""Full-scale runs of BASELINE configs C4 (2^30 keys, 10^9 queries) and C5 (3 Gbp text, 10^8 patterns of
length 20..100) on one GPU, with device-side property checks. Appends JSON lines to gpurun_out/scale.jsonl."""
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
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
# Under torchrun (one process per GPU) the index is replicated and the queries / patterns are sharded: every rank runs its
# share on its own replica, the time is the max over ranks (NCCL is used for that reduction and the barrier only).
RANK, WORLD, LOCAL = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
DEV = torch.device("cuda", LOCAL)
torch.cuda.set_device(DEV)
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=DEV)
outf = open(os.path.join(ROOT, "gpurun_out", "scale.jsonl"), "a") if RANK == 0 else None


def max_over_ranks(ms):
    if WORLD == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=DEV)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_ranks(ok):
    if WORLD == 1:
        return ok
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=DEV)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def barrier():
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()


def emit(**kw):
    if RANK != 0:
        return
    kw["n_gpus"] = WORLD
    line = json.dumps(kw)
    print(line, flush=True)
    outf.write(line + "\n")
    outf.flush()


def c4(logn=30, nq=1_000_000_000):
    L = sst.lib()
    dev = DEV
    g = torch.Generator(device=dev).manual_seed(11)
    n = 1 << logn
    keys = torch.randint(0, MAX, (n,), dtype=torch.int32, device=dev, generator=g)
    keys[0] = MAX
    keys = torch.sort(keys).values.contiguous()
    nq_total = nq
    nq = (nq + WORLD - 1) // WORLD  # this rank's contiguous share (chunk = ceil(nq / G), bench.rs:558)
    g = torch.Generator(device=dev).manual_seed(1100 + RANK)
    qs = torch.randint(0, MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
    out = torch.empty_like(qs)
    for name, build in (("stree16_left_max", lambda: sst.STree16.new_params(keys, True, False, False)),
                        ("map_b20", lambda: sst.PartitionedSTree16M.new(keys, 20))):
        t0 = time.time()
        t = build()
        torch.cuda.synchronize()
        bs = time.time() - t0
        barrier()
        ms = max_over_ranks(L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 0, 1, 3))
        torch.cuda.synchronize()
        # properties on a 10^8 slice (full-size gathers would need 8 GB more)
        sl = slice(0, min(nq, 100_000_000))
        v = out[sl]
        ok = bool((v >= qs[sl]).all())
        i = torch.searchsorted(keys, qs[:10_000_000])
        ok = ok and bool((keys[i.clamp(max=n - 1)] == out[:10_000_000]).all())
        emit(kind="c4", layout=name, logn=logn, nq=nq * WORLD, nq_per_gpu=nq, layers=t.layers(), size_mb=t.size() / 2**20, build_s=round(bs, 3), ms=ms,
             gqps=nq * WORLD / ms / 1e6, ok=all_ranks(ok))
        del t


def c5(n=3_000_000_000, npat=100_000_000):
    L = sst.lib()
    dev = DEV
    g = torch.Generator(device=dev).manual_seed(12)
    text = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev, generator=g)
    npat = (npat + WORLD - 1) // WORLD  # this rank's share of the patterns; text + SA replicated
    t0 = time.time()
    sa = sst.SaNaive.build(text)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    emit(kind="c5_build", n=n, build_s=round(build_s, 2), mem_gb=torch.cuda.mem_get_info()[0] / 2**30)
    t0 = time.time()
    viol = sa.check()
    emit(kind="c5_check", violations=viol, check_s=round(time.time() - t0, 2))
    # patterns: substrings, length uniform in [20, 100]
    g = torch.Generator(device=dev).manual_seed(1200 + RANK)
    lens = torch.randint(20, 101, (npat,), device=dev, generator=g)
    off = torch.zeros(npat + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=off[1:])
    total = int(off[-1])
    starts = torch.randint(0, n - 200, (npat,), device=dev, generator=g)
    pats = torch.empty(total + 64, dtype=torch.uint8, device=dev)
    # fill in chunks to bound the index temporaries
    CH = 5_000_000
    for a in range(0, npat, CH):
        b = min(npat, a + CH)
        ln = lens[a:b]
        o = off[a:b] - off[a]
        tot = int(off[b] - off[a])
        owner = torch.repeat_interleave(torch.arange(b - a, device=dev), ln)
        within = torch.arange(tot, device=dev) - o[owner]
        pats[int(off[a]) : int(off[b])] = text[starts[a:b][owner] + within]
        del owner, within
    lo = torch.empty(npat, dtype=torch.int32, device=dev)
    hi = torch.empty(npat, dtype=torch.int32, device=dev)
    pos = torch.empty(npat, dtype=torch.int32, device=dev)
    ref_lo = None
    variants = [("binary", sst.SA_BINARY, None), ("mlr", sst.SA_MLR, None), ("binary_nokmer", sst.SA_BINARY, "nokmer")]
    for lv in [x for x in os.environ.get("C5_SORT_LEVELS", "").split(",") if x]:  # sorted-order search at these coarse depths
        variants.append((f"binary_sorted{lv}", sst.SA_BINARY, lv))
    for name, mode, sort_lv in variants:
        sst.set_option("SA_USE_KMER", 0 if sort_lv == "nokmer" else 1)  # pivot-prefix table only: the k-mer path must agree with it
        if sort_lv is None or sort_lv == "nokmer":
            sst.set_option("SA_SORT_MIN", -1)
        else:
            sst.set_option("SA_SORT_MIN", 1)
            sst.set_option("SA_SORT_LEVELS", int(sort_lv))
        def run():
            rc = L.sst_sa_search_device(sa._h, C.c_void_p(pats.data_ptr()), C.c_void_p(off.data_ptr()), npat, mode,
                                        C.c_void_p(lo.data_ptr()), C.c_void_p(hi.data_ptr()), C.c_void_p(pos.data_ptr()), None)
            assert rc == 0, L.sst_last_error()
        run()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(a.elapsed_time(b))
        # property: the first 20 bytes of the pattern occur at the returned position
        chk = 1_000_000
        p64 = pos[:chk].long() & 0xFFFFFFFF
        got = text[(p64[:, None] + torch.arange(20, device=dev)[None, :])]
        want = pats[(off[:chk, None] + torch.arange(20, device=dev)[None, :])]
        ok = bool((got == want).all()) and bool((hi.long() > lo.long()).all())
        if ref_lo is None:
            ref_lo = lo.clone()
        same = bool((lo == ref_lo).all())
        emit(kind="c5_search", mode=name, n=n, npat=npat * WORLD, npat_per_gpu=npat, total_pattern_bytes_rank0=total, ms=ms, gpat_per_s=npat * WORLD / ms / 1e6,
             ok=all_ranks(ok), same_as_binary=all_ranks(same))


if __name__ == "__main__":
    what = sys.argv[1]
    if RANK != 0:
        sys.stdout = open(os.devnull, "w")
    if what == "c4":
        c4(int(os.environ.get("LOGN", "30")), int(os.environ.get("NQ", "1000000000")))
    elif what == "c5":
        c5(int(os.environ.get("N", "3000000000")), int(os.environ.get("NPAT", "100000000")))
