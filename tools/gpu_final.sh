#!/bin/bash
# End-of-session evidence on one GPU box: the GPU test suite, the bench line (own arm + reference arm), the ncu launch list of the
# bench command and `ncu --set full` summaries of the pipeline kernels.  usage (under gpurun): bash tools/gpu_final.sh <tag>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tag=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
( time python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2> gpurun_out/${tag}_time.log; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_time.log
python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err; echo "ref rc=$?"
LL="--steps 3 --warmup 3 --no-cpu --sa-text 0 --no-e2e --c4-log2-keys 0 --c5-text 0 --sa-rep-text 0 --c2-sizes="
python bench.py $LL > gpurun_out/${tag}_bench_launchlist_cmd.json 2>> gpurun_out/${tag}_bench.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py $LL > gpurun_out/${tag}_ncu_launchlist.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_summary.json; cut -c1-400 gpurun_out/${tag}_launch_summary.json
python tools/bk_ncu.py > gpurun_out/${tag}_bk_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:bk_ --launch-skip 8 --launch-count 4 -f -o gpurun_out/${tag}_bk_full python tools/bk_ncu.py > gpurun_out/${tag}_bk_ncu.log 2>&1
tail -2 gpurun_out/${tag}_bk_ncu.log
