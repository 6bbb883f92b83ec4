#!/bin/bash
# A/B helper for the reordered-batch pipeline on the GPU box: the pipeline's parity tests, then warm per-stage times.
# usage (under gpurun): bash tools/gpu_ab.sh <tag> [ENV=VAL ...]   -> gpurun_out/<tag>_{tests.log,stages.jsonl}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tag=$1; shift
python -m pytest tests/test_gpu_bucketed.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/${tag}_tests.log
env "$@" TAG=$tag python tools/bucketed_stages.py >> gpurun_out/${tag}_stages.jsonl 2>> gpurun_out/${tag}_stages.err
env "$@" python tools/bucketed_once.py >> gpurun_out/${tag}_stages.jsonl 2>> gpurun_out/${tag}_stages.err
cat gpurun_out/${tag}_tests.log gpurun_out/${tag}_stages.jsonl
