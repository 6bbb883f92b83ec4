cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r2_s5_pytest.log 2>&1; cat gpurun_out/r2_s5_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py > gpurun_out/r2_s5_bench.json 2> gpurun_out/r2_s5_bench.err ) 2>&1 | tail -3; echo rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_s5_bench_ref.json 2>> gpurun_out/r2_s5_bench.err
