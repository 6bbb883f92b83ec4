#!/bin/bash
# Local helper: rebuild the library and only then spend a gpurun call.  usage: tools/grun.sh [--gpus N] <timeout> '<command>'
set -e
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
make -C /root/repo/suffix-array-searching_b200/csrc -j8 -s 2>&1 | grep -E "error|Error" && { echo BUILD FAILED; exit 1; }
make -C /root/repo/suffix-array-searching_b200/csrc -j8 -s >/dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
/usr/local/graft/bin/gpurun $G --timeout $1 -- "$2" 2>&1 | tail -${TAIL:-12}
