import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "suffix-array-searching_b200"))
import sst_b200 as sst
L = sst.lib()
for b in (64 << 20, 1 << 30):
    print(b, L.sst_probe_gather64(0, b, 100_000_000, 2, 1) / 64, flush=True)
