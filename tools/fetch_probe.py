import os, sys, subprocess
ROOT='/root/repo' if os.path.exists('/root/repo/tools') else os.getcwd()
code = r'''
import os,sys
sys.path.insert(0, os.path.join(os.getcwd(), "suffix-array-searching_b200"))
import sst_b200 as sst
L=sst.lib()
for b in (64<<20, 1<<30, 8<<30):
    print(os.environ.get("SST_L2_FETCH"), b, round(L.sst_probe_gather64(0, b, 200000000, 2, 3)/64,2), "Gnodes/s", flush=True)
'''
for f in ("0","32","64","128"):
    env=dict(os.environ, SST_L2_FETCH=f, SST_DEBUG="1")
    subprocess.run([sys.executable,"-c",code],env=env)
