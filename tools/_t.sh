cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_bucketed.py tests/test_gpu_stree.py -x -q -m gpu 2>&1 | tail -8
python - <<'PY'
import sys, ctypes as C, json
sys.path.insert(0, "suffix-array-searching_b200")
import torch, sst_b200 as sst
L = sst.lib(); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
n = 1 << 28; nq = 100_000_000
keys = torch.randint(0, sst.MAX, (n,), dtype=torch.int32, device=dev, generator=g); keys[0] = sst.MAX
keys = torch.sort(keys).values.contiguous()
qs = torch.randint(0, sst.MAX, (nq,), dtype=torch.int32, device=dev, generator=g)
out = torch.empty_like(qs)
ref = sst.STree16.new_params(keys, True, False, False).query(qs)
for name in ("PartitionedSTree16C", "PartitionedSTree16"):
    t = getattr(sst, name).new(keys, 20)
    ms = L.sst_time_query_device(t._h, C.c_void_p(qs.data_ptr()), nq, C.c_void_p(out.data_ptr()), None, 0, 2, 5)
    sch, ln = C.c_int(0), C.c_int(0); L.sst_query_plan(t._h, nq, 0, 0, C.byref(sch), C.byref(ln))
    print(json.dumps({"layout": name, "b": 20, "ms": round(ms, 3), "gqps": round(nq / ms / 1e6, 2), "scheme": sch.value, "equal_plain": bool((out == ref).all()), "size_mb": t.size() >> 20}))
    del t
PY
