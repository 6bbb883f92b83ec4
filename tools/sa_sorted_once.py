"""One sorted-order SA search at C3 scale (for an ncu launch list)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import torch
import sst_b200 as sst
L = sst.lib(); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
n, npat, plen = 100_000_000, 10_000_000, 32
text = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev, generator=g)
sa = sst.SaNaive.build(text)
starts = torch.randint(0, n - 200, (npat,), device=dev, generator=g)
pats = text[(starts[:, None] + torch.arange(plen, device=dev)[None, :]).reshape(-1)].contiguous()
off = (torch.arange(npat + 1, device=dev, dtype=torch.int64) * plen).contiguous()
lo = torch.empty(npat, dtype=torch.int32, device=dev); hi = torch.empty_like(lo); pos = torch.empty_like(lo)
for _ in range(2):
    rc = L.sst_sa_search_device(sa._h, C.c_void_p(pats.data_ptr()), C.c_void_p(off.data_ptr()), npat, 0, C.c_void_p(lo.data_ptr()), C.c_void_p(hi.data_ptr()), C.c_void_p(pos.data_ptr()), None)
    assert rc == 0
torch.cuda.synchronize()
print("ok")
