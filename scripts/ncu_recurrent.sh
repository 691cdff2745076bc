#!/bin/bash
# ncu --set full capture (with source-level stall sampling) of the recurrence kernel for one geometry.
# usage: scripts/ncu_recurrent.sh <geom|auto> <tag> [T]
set -u
OUT=gpurun_out
mkdir -p $OUT
g=${1:-auto}; tag=${2:-rc}; T=${3:-128}
if [ "$g" = auto ]; then unset HSSB_RC_GEOM; else export HSSB_RC_GEOM=$g; fi
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:tc_recurrent -s 2 -c 2 -f -o $OUT/ncu_$tag \
    python scripts/trace_recurrent.py 512 $T 8 > $OUT/ncu_$tag.log 2>&1
echo "ncu $tag rc=$?"; tail -3 $OUT/ncu_$tag.log
