#!/bin/bash
# ncu --set full capture of kernels matching a regex during one BiLSTM forward (B=512).
# usage: scripts/ncu_kernel.sh <regex> <tag> [T] [skip] [count]
set -u
OUT=gpurun_out
mkdir -p $OUT
re=${1:-tc_inproj}; tag=${2:-k4}; T=${3:-2000}; skip=${4:-2}; cnt=${5:-2}
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c $cnt -f -o $OUT/ncu_$tag \
    python scripts/trace_recurrent.py 512 $T 8 > $OUT/ncu_$tag.log 2>&1
echo "ncu $tag rc=$?"; tail -3 $OUT/ncu_$tag.log
