#!/bin/bash
# ncu --set full (with source) of the SA search kernel at the C3 shape (or N/NPAT/LMIN/LMAX from the environment).
# usage (under gpurun): bash tools/gpu_sa_ncu.sh <tag> [ENV=VAL ...]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tag=$1; shift
[ -n "$NOPLAIN" ] || env "$@" TAG=$tag python tools/sa_bench.py > gpurun_out/${tag}_sa.jsonl 2> gpurun_out/${tag}_sa.err || { tail -5 gpurun_out/${tag}_sa.err; exit 1; }
cat gpurun_out/${tag}_sa.jsonl
env "$@" ncu --set full --import-source on --clock-control none -k regex:sa_search_thread_kernel --launch-skip ${SKIP:-17} --launch-count 1 -f -o gpurun_out/${tag}_sa python tools/sa_bench.py > gpurun_out/${tag}_sa_ncu.log 2>&1
tail -2 gpurun_out/${tag}_sa_ncu.log
