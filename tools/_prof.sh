cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_s2_bench.json 2> gpurun_out/r2_s2_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_s2_bench_ref.json 2>> gpurun_out/r2_s2_bench.err
ARGS="--steps 3 --warmup 3 --no-cpu --sa-text 0 --no-e2e --c4-log2-keys 0 --c5-text 0"
python bench.py $ARGS > gpurun_out/r2_s2_bench_short.json 2>> gpurun_out/r2_s2_bench.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_s2_launches.csv python bench.py $ARGS > gpurun_out/r2_s2_ncu_bench.log 2>&1
python tools/bk_ncu.py > gpurun_out/r2_s2_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:'bk_(part|items|search2|unperm)' --launch-skip 8 --launch-count 4 -f -o gpurun_out/r2_s2_bk_full python tools/bk_ncu.py > gpurun_out/r2_s2_ncu_bk.log 2>&1
tail -2 gpurun_out/r2_s2_ncu_bk.log
cut -c1-400 gpurun_out/r2_s2_bench.json
