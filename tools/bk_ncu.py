"""Three runs of the reordered-batch pipeline at LOGN keys / NQ queries (for ncu: profile the kernels of the last run)."""
import ctypes as C, os, sys
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
for _ in range(int(os.environ.get("RUNS", "3"))):
    rc = L.sst_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 7, st)
    assert rc == 0, L.sst_last_error()
torch.cuda.synchronize()
i = torch.searchsorted(keys, qs[:1000000])
print("ok", bool((keys[i.clamp(max=n - 1)] == out[:1000000]).all()))
