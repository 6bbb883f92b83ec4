import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "suffix-array-searching_b200"))
import sst_b200 as sst
L = sst.lib()
for mode in (1000, 1064, 1128, 1256):
    for b in (1 << 30,):
        print(mode, b, round(L.sst_probe_gather64(0, b, 100_000_000, mode, 2) / 64, 2), "Gnodes/s", flush=True)
