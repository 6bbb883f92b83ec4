"""Reordered-batch pipeline (SCHEME_BUCKETED) vs the rank-table kernel at LOGN keys / NQ queries: equality of
results and CUDA-event times; the option BK_TIMING=2 prints the per-stage split."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import torch
import sst_b200 as sst
L = sst.lib(); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
n = 1 << int(os.environ.get("LOGN", "28"))
nq = int(os.environ.get("NQ", "100000000"))
keys = torch.randint(0, sst.MAX, (n,), dtype=torch.int32, device=dev, generator=g); keys[0] = sst.MAX
keys = torch.sort(keys).values.contiguous()
t = sst.STree16.new_params(keys, True, False, False)
qs = torch.randint(0, sst.MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
if os.environ.get("SKEW") == "equal":      # every query the same key: one bucket, every lane of a warp collides
    qs[:] = int(keys[n // 3])
elif os.environ.get("SKEW") == "narrow":   # all queries inside one bucket
    qs = (int(keys[n // 3]) + (qs % 100_000)).to(torch.int32).contiguous()
elif os.environ.get("SKEW") == "sorted":
    qs = torch.sort(qs).values.contiguous()
v1, i1 = t.query(qs, sst.SCHEME_TABLE, want_index=True)
v2, i2 = t.query(qs, sst.SCHEME_BUCKETED, want_index=True)
torch.cuda.synchronize()
print(json.dumps({"values_equal": bool((v1 == v2).all()), "indices_equal": bool((i1 == i2).all())}), flush=True)
del i1, i2, v2
out = torch.empty_like(qs)
for name, scheme in (("table", 5), ("bucketed", 7)):
    ms = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, scheme, 2, 5)
    print(json.dumps({"scheme": name, "ms": round(ms, 3), "gqps": round(nq / ms / 1e6, 2), "ok": bool((out == v1).all())}), flush=True)
