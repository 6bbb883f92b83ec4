cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_sa.py tests/test_gpu_bench.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15
python bench.py --no-cpu --c4-log2-keys 0 --c5-text 0 --no-e2e --steps 3 > gpurun_out/r2_t30_bench_rep.json 2> gpurun_out/r2_t30_bench_rep.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_t30_bench_rep.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('sa_repetitive'))[:1500])
print({k:v for k,v in d['sa'].items() if 'patterns_per_s' in k})
PY
