cd $GRAFT_REPO_ROOT
python tools/sanitize_smoke.py 2>&1 | tail -3
