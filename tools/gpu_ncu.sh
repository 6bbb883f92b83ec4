#!/bin/bash
# One `ncu --set full` capture (with source) of a pipeline kernel on the GPU box, after the plain run has exited 0.
# usage (under gpurun): bash tools/gpu_ncu.sh <tag> <kernel-regex> [ENV=VAL ...]   -> gpurun_out/<tag>.ncu-rep
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tag=$1; kern=$2; shift 2
env "$@" python tools/bk_ncu.py > gpurun_out/${tag}_plain.log 2>&1 || { cat gpurun_out/${tag}_plain.log; exit 1; }
env "$@" ncu --set full --import-source on --clock-control none -k regex:$kern --launch-skip 2 --launch-count 1 -f -o gpurun_out/${tag} python tools/bk_ncu.py > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
