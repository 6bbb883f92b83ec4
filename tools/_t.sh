cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_sa.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -8
TAG=cells python tools/sa_bench.py 2>&1 | tail -2
SST_SA_USE_CELLS=0 TAG=nocells python tools/sa_bench.py 2>&1 | tail -2
