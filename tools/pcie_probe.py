"""PCIe ceiling on this box: pinned H2D alone, D2H alone, both at once (two streams), 400 MB each, CUDA-event timed."""
import json, torch
dev = torch.device("cuda", 0)
n = 100_000_000
h_in = torch.empty(n, dtype=torch.int32).pin_memory(); h_out = torch.empty(n, dtype=torch.int32).pin_memory()
d_in = torch.empty(n, dtype=torch.int32, device=dev); d_out = torch.zeros(n, dtype=torch.int32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a); s2.wait_event(a)
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        e1, e2 = torch.cuda.Event(), torch.cuda.Event()
        e1.record(s1); e2.record(s2)
        torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
r = {"h2d_ms": run(True, False), "d2h_ms": run(False, True), "both_ms": run(True, True)}
r["h2d_gbs"] = 0.4 / r["h2d_ms"] * 1e3; r["d2h_gbs"] = 0.4 / r["d2h_ms"] * 1e3; r["both_gbs_each_way"] = 0.4 / r["both_ms"] * 1e3
r["ceiling_gqps"] = n / r["both_ms"] / 1e6
print(json.dumps(r))
# chunked pipeline without a kernel: H2D chunk c on s1 -> event -> D2H of the same chunk on s2
import time
for chunk in (1 << 20, 1 << 22, 1 << 23):
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for c in range(0, n, chunk):
            e = min(n, c + chunk)
            with torch.cuda.stream(s1):
                d_in[c:e].copy_(h_in[c:e], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(s1)
            with torch.cuda.stream(s2):
                s2.wait_event(ev)
                h_out[c:e].copy_(d_in[c:e], non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(json.dumps({"chunked_copy_only": chunk, "ms": round(best * 1e3, 3), "gqps": round(n / best / 1e9, 2)}))
