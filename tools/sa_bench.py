"""SA search timing at config C3 / C5 shape on one GPU: N text bytes, NPAT patterns of length PLEN (or LMIN..LMAX), binary and mlr.
Prints one JSON line per mode; SST_B200_LIB selects an alternative build of the library (A/B runs)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import torch
import sst_b200 as sst
L = sst.lib(); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
n, npat = int(os.environ.get("N", "100000000")), int(os.environ.get("NPAT", "10000000"))
lmin, lmax = int(os.environ.get("LMIN", "32")), int(os.environ.get("LMAX", "32"))
text = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev, generator=g)
sa = sst.SaNaive.build(text)
lens = torch.randint(lmin, lmax + 1, (npat,), device=dev, generator=g)
off = torch.zeros(npat + 1, dtype=torch.int64, device=dev); torch.cumsum(lens, 0, out=off[1:])
total = int(off[-1])
starts = torch.randint(0, n - 200, (npat,), device=dev, generator=g)
pats = torch.empty(total + 64, dtype=torch.uint8, device=dev)
CH = 5_000_000
for a in range(0, npat, CH):
    b = min(npat, a + CH)
    owner = torch.repeat_interleave(torch.arange(b - a, device=dev), lens[a:b])
    within = torch.arange(int(off[b] - off[a]), device=dev) - (off[a:b] - off[a])[owner]
    pats[int(off[a]):int(off[b])] = text[starts[a:b][owner] + within]
    del owner, within
lo = torch.empty(npat, dtype=torch.int32, device=dev); hi = torch.empty_like(lo); pos = torch.empty_like(lo)
ref = None
for name, mode in (("binary", 0), ("mlr", 1)):
    def run():
        rc = L.sst_sa_search_device(sa._h, C.c_void_p(pats.data_ptr()), C.c_void_p(off.data_ptr()), npat, mode, C.c_void_p(lo.data_ptr()),
                                    C.c_void_p(hi.data_ptr()), C.c_void_p(pos.data_ptr()), None)
        assert rc == 0, L.sst_last_error()
    run(); torch.cuda.synchronize()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record()
    for _ in range(3): run()
    b_.record(); torch.cuda.synchronize()
    ms = a_.elapsed_time(b_) / 3
    if ref is None: ref = (lo.clone(), hi.clone())
    print(json.dumps({"tag": os.environ.get("TAG", ""), "lib": os.path.basename(sst.LIB_PATH), "n": n, "npat": npat, "len": [lmin, lmax], "mode": name, "ms": round(ms, 4),
                      "gpat_per_s": round(npat / ms / 1e6, 3), "same_as_binary": bool((lo == ref[0]).all() and (hi == ref[1]).all()),
                      "lo_sum": int(lo.long().sum()), "hi_sum": int(hi.long().sum())}), flush=True)
