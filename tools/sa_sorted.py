"""SA search at config C3 scale: patterns in the caller's order (SST_SA_SORT_MIN huge) vs sorted-order search, for a
few coarse depths; checks that every output is identical."""
import ctypes as C, os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import torch
import sst_b200 as sst
L = sst.lib(); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
n = int(os.environ.get("N", "100000000")); npat = int(os.environ.get("NPAT", "10000000")); plen = 32
text = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev, generator=g)
sa = sst.SaNaive.build(text)
starts = torch.randint(0, n - 200, (npat,), device=dev, generator=g)
pats = text[(starts[:, None] + torch.arange(plen, device=dev)[None, :]).reshape(-1)].contiguous()
off = (torch.arange(npat + 1, device=dev, dtype=torch.int64) * plen).contiguous()
def run(mode):
    lo = torch.empty(npat, dtype=torch.int32, device=dev); hi = torch.empty_like(lo); pos = torch.empty_like(lo)
    def once():
        rc = L.sst_sa_search_device(sa._h, C.c_void_p(pats.data_ptr()), C.c_void_p(off.data_ptr()), npat, mode, C.c_void_p(lo.data_ptr()), C.c_void_p(hi.data_ptr()), C.c_void_p(pos.data_ptr()), None)
        assert rc == 0, L.sst_last_error()
    once(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): once()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 3, (lo, hi, pos)
os.environ["SST_SA_SORT_MIN"] = str(1 << 62)
base = {}
for mode in (0, 1):
    ms, base[mode] = run(mode)
    print(json.dumps({"order": "caller", "mode": mode, "ms": round(ms, 3), "gpat_s": round(npat / ms / 1e6, 3)}), flush=True)
os.environ["SST_SA_SORT_MIN"] = "1"
for lv in ([] if os.environ.get("SKIP_SORTED") else [12, 15, 18, 21, 24]):
    os.environ["SST_SA_SORT_LEVELS"] = str(lv)
    for mode in (0, 1):
        ms, out = run(mode)
        same = all(bool((x == y).all()) for x, y in zip(out, base[mode]))
        print(json.dumps({"order": "sorted", "coarse_levels": lv, "mode": mode, "ms": round(ms, 3), "gpat_s": round(npat / ms / 1e6, 3), "identical": same}), flush=True)
