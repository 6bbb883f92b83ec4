cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2_t27_pytest.log 2>&1
cat gpurun_out/r2_t27_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
