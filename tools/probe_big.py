"""Random 64-byte gather rate (sst_probe_gather64) over footprints of 1 .. 144 GiB: the TLB reach of the part."""
import ctypes as C, json, os, sys
sys.path.insert(0, "suffix-array-searching_b200")
import torch, sst_b200 as sst
L = sst.lib()
L.sst_probe_gather64.restype = C.c_double
L.sst_probe_gather64.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int]
for gb in [int(x) for x in os.environ.get("GIBS", "1,4,16,48,96,144").split(",")]:
    r = L.sst_probe_gather64(0, gb << 30, 100_000_000, 2, 3)
    print(json.dumps({"kind": "probe", "gib": gb, "gbs": round(r, 1), "ggathers_per_s": round(r / 64, 2)}), flush=True)
