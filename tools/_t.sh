cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_tools.py tests/test_formats.py -x -q -m gpu 2>&1 | tail -15
