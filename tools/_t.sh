cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_bench.py -x -q -m gpu 2>&1 | tail -5
python bench.py --no-cpu --c4-log2-keys 0 --c5-text 0 --no-e2e --steps 3 --sa-text 0 --sa-rep-text 0 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['c2'])[:1500])"
