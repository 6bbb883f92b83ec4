"""Host-buffer path (sst_query: pinned host -> H2D -> kernel -> D2H) at several chunk sizes; 2^28 keys, 10^8 queries."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import torch, sst_b200 as sst
L = sst.lib(); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
n = 1 << int(os.environ.get("LOGN", "28")); nq = int(os.environ.get("NQ", "100000000"))
keys = torch.randint(0, sst.MAX, (n,), dtype=torch.int32, device=dev, generator=g); keys[0] = sst.MAX
keys = torch.sort(keys).values.contiguous()
t = sst.STree16.new_params(keys, True, False, False)
qs = torch.randint(0, sst.MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
hq = torch.empty(nq, dtype=torch.int32).pin_memory(); hv = torch.empty(nq, dtype=torch.int32).pin_memory()
hq.copy_(qs); torch.cuda.synchronize()
want = t.query(qs)
for var in sys.argv[1:] or ["SST_CHUNK=4194304"]:
    kv = dict(x.split("=") for x in var.split(","))
    for k, v in kv.items(): sst.set_option(k, int(v))
    hv.zero_()
    def step():
        rc = L.sst_query(t._h, C.c_void_p(hq.data_ptr()), nq, C.c_void_p(hv.data_ptr()), None, 0)
        assert rc == 0, L.sst_last_error()
    step(); step()
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); step(); best = min(best, time.perf_counter() - t0)
    ok = bool((hv.to(dev) == want).all())
    print(json.dumps({"var": var, "ms": round(best * 1e3, 3), "gqps": round(nq / best / 1e9, 2), "ok": ok}), flush=True)
    sst.reset_options()
