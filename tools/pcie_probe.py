"""PCIe ceiling on this box: pinned H2D alone, D2H alone, both at once (two streams), 400 MB each, CUDA-event timed.
Under torchrun (one process per GPU) every rank copies at the same time between barriers and the line reports the
max-over-ranks time and the aggregate GB/s over all GPUs: the copy-only ceiling of the host-buffer path (sst_query)."""
import json, os, time, torch
RANK, WORLD, LOCAL = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", LOCAL)
torch.cuda.set_device(dev)
dist = None
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(x):
    if dist is None:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


n = 100_000_000
h_in = torch.empty(n, dtype=torch.int32).pin_memory(); h_out = torch.empty(n, dtype=torch.int32).pin_memory()
d_in = torch.empty(n, dtype=torch.int32, device=dev); d_out = torch.zeros(n, dtype=torch.int32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    best = 1e9
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        best = min(best, max_over_ranks((time.perf_counter() - t0) * 1e3))
    return best


r = {"n_gpus": WORLD, "h2d_ms": run(True, False), "d2h_ms": run(False, True), "both_ms": run(True, True)}
r["h2d_gbs_aggregate"] = WORLD * 0.4 / r["h2d_ms"] * 1e3; r["d2h_gbs_aggregate"] = WORLD * 0.4 / r["d2h_ms"] * 1e3
r["both_gbs_each_way_aggregate"] = WORLD * 0.4 / r["both_ms"] * 1e3
r["ceiling_gqps_aggregate"] = WORLD * n / r["both_ms"] / 1e6
if RANK == 0:
    print(json.dumps(r), flush=True)
# chunked pipeline without a kernel: H2D chunk c on s1 -> event -> D2H of the same chunk on s2 (the shape of sst_query)
for chunk in (1 << 22,):
    best = 1e9
    for _ in range(5):
        barrier(); t0 = time.perf_counter()
        for c in range(0, n, chunk):
            e = min(n, c + chunk)
            with torch.cuda.stream(s1):
                d_in[c:e].copy_(h_in[c:e], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(s1)
            with torch.cuda.stream(s2):
                s2.wait_event(ev)
                h_out[c:e].copy_(d_in[c:e], non_blocking=True)
        torch.cuda.synchronize(); best = min(best, max_over_ranks(time.perf_counter() - t0))
    if RANK == 0:
        print(json.dumps({"n_gpus": WORLD, "chunked_copy_only": chunk, "ms": round(best * 1e3, 3), "gqps_aggregate": round(WORLD * n / best / 1e9, 2)}), flush=True)
if dist is not None:
    dist.destroy_process_group()
