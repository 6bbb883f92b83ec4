"""Warm per-stage times of the reordered-batch pipeline (mean of 5 runs with the option BK_TIMING=1) at LOGN keys / NQ queries."""
import os, sys, ctypes as C, json
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
out = torch.empty_like(qs)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3): L.sst_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 7, st)
ms_all = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 7, 1, 5)
sst.set_option("BK_TIMING", 1)
acc = [0.0] * 5
for _ in range(5):
    L.sst_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 7, st)
    torch.cuda.synchronize()
    ms = (C.c_double * 5)(); L.sst_last_stage_ms(ms, 5)
    acc = [a + b for a, b in zip(acc, ms)]
print(json.dumps({"tag": os.environ.get("TAG", ""), "ms": round(ms_all, 3), "gqps": round(nq / ms_all / 1e6, 2),
                  "stages_ms": dict(zip(["partition", "plan", "-", "search", "unpermute"], [round(a / 5, 3) for a in acc]))}))
