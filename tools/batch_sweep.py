"""Throughput vs batch size at 2^28 keys: table kernel (TMA-staged rank table, 180 KB per CTA) vs group kernel."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import torch
import sst_b200 as sst
L = sst.lib(); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
n = 1 << int(os.environ.get("LOGN", "28"))
keys = torch.randint(0, sst.MAX, (n,), dtype=torch.int32, device=dev, generator=g); keys[0] = sst.MAX
keys = torch.sort(keys).values.contiguous()
t = sst.STree16.new_params(keys, True, False, False)
qs = torch.randint(0, sst.MAX, (1 << 26,), dtype=torch.int32, device=dev, generator=g)
out = torch.empty_like(qs)
for lg in range(10, 27, 2):
    nq = 1 << lg
    row = {"log2_nq": lg}
    for name, scheme in (("table", 5), ("group2", 3), ("generic", 4)):
        ms = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, scheme, 3, 20 if lg < 22 else 5)
        row[name + "_us"] = round(ms * 1e3, 1)
        row[name + "_gqps"] = round(nq / ms / 1e6, 2)
    print(json.dumps(row), flush=True)
