cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_multi.py tests/test_gpu_sa.py tests/test_gpu_stree.py -x -q -m gpu 2>&1 | tail -15
