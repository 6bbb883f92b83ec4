"""In-process multi-GPU check and timing: sst_multi_* with one worker thread + streams per device."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import sst_b200 as sst
from oracle import oracle as O
nd = sst.device_count()
print("devices", nd)
rng = np.random.default_rng(1)
n = 1 << 26
vals = np.sort(rng.integers(0, sst.MAX, n, dtype=np.uint32)); vals[-1] = sst.MAX
nq = 64_000_000
q = sst.PinnedArray(nq); q.array[:] = rng.integers(0, sst.MAX, nq, dtype=np.uint32)
out = sst.PinnedArray(nq)
import ctypes as C
L = sst.lib()
ev, _ = O.lower_bound(vals, q.array[:200_000])
for g in sorted({1, min(2, nd), min(4, nd), nd}):
    m = sst.MultiIndex.stree(vals, list(range(g)), left_max=True)
    for it in range(3):
        t0 = time.perf_counter()
        rc = L.sst_multi_query(m._h, q.array.ctypes.data_as(C.c_void_p), nq, out.array.ctypes.data_as(C.c_void_p), None, 0)
        dt = time.perf_counter() - t0
        assert rc == 0, L.sst_last_error()
    assert (out.array[:200_000] == ev).all() and (out.array[-1000:] == O.lower_bound(vals, q.array[-1000:])[0]).all()
    print(f"gpus={g} e2e {nq/dt/1e9:.2f} Gq/s ({dt*1e3:.1f} ms)", flush=True)
    del m
