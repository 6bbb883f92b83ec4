cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_bucketed.py tests/test_gpu_stree.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -6
python - <<'PY'
import sys, ctypes as C, json
sys.path.insert(0, "suffix-array-searching_b200")
import torch, sst_b200 as sst
L = sst.lib(); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
n = 1 << 30; nq = 1_000_000_000
keys = torch.randint(0, sst.MAX, (n,), dtype=torch.int32, device=dev, generator=g); keys[0] = sst.MAX
keys = torch.sort(keys).values.contiguous()
torch.cuda.empty_cache()
qs = torch.randint(0, sst.MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
out = torch.empty_like(qs)
t = sst.STree16.new_params(keys, True, False, False)
for sub in (27, 28, 29, 30):
    sst.set_option("BK_SUB_LOG2", sub)
    ms = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 0, 1, 3)
    i = torch.searchsorted(keys, qs[:2000000])
    print(json.dumps({"sub_log2": sub, "ms": round(ms, 2), "gqps": round(nq / ms / 1e6, 2), "ok": bool((keys[i.clamp(max=n-1)] == out[:2000000]).all())}), flush=True)
PY
