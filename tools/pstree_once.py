"""One partitioned-layout query launch at 2^28 keys / 10^8 queries (for ncu). LAYOUT=map|compact|l1|overlap|simple"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import torch
import sst_b200 as sst
L = sst.lib(); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
n, nq = 1 << int(os.environ.get("LOGN", "28")), 100_000_000
keys = torch.randint(0, sst.MAX, (n,), dtype=torch.int32, device=dev, generator=g); keys[0] = sst.MAX
keys = torch.sort(keys).values.contiguous()
qs = torch.randint(0, sst.MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
out = torch.empty_like(qs)
cls = {"map": sst.PartitionedSTree16M, "compact": sst.PartitionedSTree16C, "l1": sst.PartitionedSTree16L, "overlap": sst.PartitionedSTree16O, "simple": sst.PartitionedSTree16}
for name in os.environ.get("LAYOUT", "compact,simple").split(","):
    t = cls[name].new(keys, 20)
    ms = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 0, 1, 2)
    print(name, round(nq / ms / 1e6, 2), "Gq/s", t.layers(), "layers", round(t.size() / 2**20), "MB", flush=True)
    del t
