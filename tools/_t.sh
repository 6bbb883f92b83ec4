cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_bucketed.py tests/test_gpu_host.py -x -q -m gpu 2>&1 | tail -15
