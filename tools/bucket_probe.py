"""Upper bound for a reordered batch: rate of the search kernels when the 10^8 queries arrive bucketed by
their top key bits (stable), i.e. what the search phase of a partition -> search -> un-permute pipeline
would run at.  LOGN keys (default 2^28)."""
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
qs0 = torch.randint(0, sst.MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
out = torch.empty_like(qs0)
for bits in (0, 3, 4, 5, 6, 8, 10, 12, 16, 31):
    if bits == 0: qs = qs0
    elif bits == 31: qs = torch.sort(qs0).values
    else:
        b = (qs0 >> (31 - bits)).to(torch.int16 if bits < 15 else torch.int32)
        qs = qs0[torch.sort(b, stable=True).indices].contiguous()
        del b
    torch.cuda.synchronize()
    row = {"bucket_bits": bits}
    for name, scheme in (("table", 5), ("group2", 3), ("generic", 4)):
        ms = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, scheme, 2, 5)
        row[name + "_ms"] = round(ms, 3)
        row[name + "_gqps"] = round(nq / ms / 1e6, 2)
    print(json.dumps(row), flush=True)
    del qs
