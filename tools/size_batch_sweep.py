"""Tree size x batch size sweep on one GPU: the rank-table kernel vs the reordered-batch pipeline vs what SCHEME_AUTO picks.
One JSON line per (log2 keys, log2 queries).  Used to place AUTO's crossover (csrc/stree_search.cu: resolve_scheme)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import torch
import sst_b200 as sst
L = sst.lib(); dev = torch.device("cuda", 0)
logns = [int(x) for x in os.environ.get("LOGNS", "25,26,27,28,29,30").split(",")]
lognqs = [int(x) for x in os.environ.get("LOGNQS", "18,20,21,22,23,24,25,26,27").split(",")]
g = torch.Generator(device=dev).manual_seed(3)
qs = torch.randint(0, sst.MAX, (1 << max(lognqs),), dtype=torch.int32, device=dev, generator=g)
out = torch.empty_like(qs)
for logn in logns:
    n = 1 << logn
    keys = torch.randint(0, sst.MAX, (n,), dtype=torch.int32, device=dev, generator=g); keys[0] = sst.MAX
    keys = torch.sort(keys).values.contiguous()
    torch.cuda.empty_cache()
    t = sst.STree16.new_params(keys, True, False, False)
    for lg in lognqs:
        nq = 1 << lg
        row = {"log2_keys": logn, "log2_nq": lg}
        for name, scheme in (("table", 5), ("bucketed", 7), ("auto", 0)):
            iters = 20 if lg < 22 else 5
            ms = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, scheme, 3, iters)
            row[name + "_us"] = round(ms * 1e3, 1)
            row[name + "_gqps"] = round(nq / ms / 1e6, 2)
        sch, ln = C.c_int(0), C.c_int(0)
        L.sst_query_plan(t._h, nq, 0, 0, C.byref(sch), C.byref(ln))
        row["auto_scheme"] = sch.value
        row["auto_vs_best"] = round(row["auto_gqps"] / max(row["table_gqps"], row["bucketed_gqps"]), 3)
        print(json.dumps(row), flush=True)
    del t, keys
