"""A tree of N keys (default 2^31: more slots than 2^30, the 16-bit separators with 16 keys each) through the reordered-batch pipeline:
exact parity of a sample against torch.searchsorted, agreement with the direct kernel on the whole batch, stage times."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import torch, sst_b200 as sst
L = sst.lib(); dev = torch.device("cuda", 0)
n = int(os.environ.get("N", str(1 << 31))); nq = int(os.environ.get("NQ", str(1 << 27)))
g = torch.Generator(device=dev).manual_seed(3)
# sorted by construction (a sort of 2^31 keys needs three copies): key i = floor(i * MAX / n) plus a few duplicates and gaps
keys = torch.empty(n, dtype=torch.int32, device=dev)
CH = 1 << 28
for a in range(0, n, CH):
    b = min(n, a + CH)
    i = torch.arange(a, b, dtype=torch.int64, device=dev)
    k = (i * sst.MAX) // n
    k = k - (k % 3 == 1).long()          # every third value collapses onto its predecessor: duplicates and gaps, still sorted
    keys[a:b] = k.to(torch.int32)
    del i, k
keys[-1] = sst.MAX
t = sst.STree16.new_params(keys, True, False, False)
qs = torch.randint(0, sst.MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
out = torch.empty_like(qs); out2 = torch.empty_like(qs)
idx = torch.empty(nq, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
sch, ln = C.c_int(0), C.c_int(0)
L.sst_query_plan(t._h, nq, 0, 0, C.byref(sch), C.byref(ln))
rc = L.sst_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), C.c_void_p(idx.data_ptr()), 7, st)
assert rc == 0, L.sst_last_error()
rc = L.sst_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out2.data_ptr()), None, 5, st)
assert rc == 0, L.sst_last_error()
torch.cuda.synchronize()
m = 1 << 22
i = torch.searchsorted(keys, qs[:m])
ok_val = bool((keys[i.clamp(max=n - 1)] == out[:m]).all())
ok_idx = bool((idx[:m] == i).all())
same = bool((out == out2).all())
ms_b = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 7, 2, 5)
ms_t = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 5, 2, 5)
ms_a = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 0, 2, 5)
print(json.dumps({"tag": os.environ.get("TAG", ""), "n_keys": n, "nq": nq, "auto_scheme": sch.value, "values_ok": ok_val, "index_ok": ok_idx, "pipeline_equals_direct": same,
                  "pipeline_gqps": round(nq / ms_b / 1e6, 2), "direct_gqps": round(nq / ms_t / 1e6, 2), "auto_gqps": round(nq / ms_a / 1e6, 2)}))
