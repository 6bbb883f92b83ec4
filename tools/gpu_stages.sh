#!/bin/bash
# Warm per-stage times of the pipeline for several option settings in one GPU call.
# usage (under gpurun): bash tools/gpu_stages.sh <tag> "ENV=VAL ..." "ENV=VAL ..." ...   (one stage line per setting; "" = defaults)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tag=$1; shift
for cfg in "$@"; do
  env $cfg TAG="$tag $cfg" python tools/bucketed_stages.py >> gpurun_out/${tag}_stages.jsonl 2>> gpurun_out/${tag}_stages.err
done
cat gpurun_out/${tag}_stages.jsonl
