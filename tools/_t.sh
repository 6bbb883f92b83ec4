cd $GRAFT_REPO_ROOT
for mb in 0 4 5 0 5; do SST_SA_MINB=$mb TAG=minb$mb python tools/sa_bench.py 2>&1 | tail -2 | head -1 | cut -c1-160; done
