#!/bin/bash
# K5 experiments: parity of every recurrence geometry, then clock64 traces + timings per geometry.
# usage: scripts/gpu_k5.sh "geom1 geom2 ..."   (HSSB_RC_GEOM values; "auto" = library default)
set -u
OUT=gpurun_out
mkdir -p $OUT
GEOMS="${1:-auto 64,2,1 64,2,2}"
run() { # name timeout cmd...
    local name=$1 to=$2; shift 2
    echo "=== $name ($(date +%T))" | tee -a $OUT/summary.txt
    timeout -s KILL $to "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "rc=$rc" | tee -a $OUT/summary.txt
    tail -n ${TAIL:-12} $OUT/$name.log | tee -a $OUT/summary.txt
    return $rc
}
run geoms 600 python -m pytest tests/test_lstm_gpu.py -q -m gpu -x -k "geometries"
for g in $GEOMS; do
    n=$(echo $g | tr ',' '_')
    if [ "$g" = auto ]; then unset HSSB_RC_GEOM; else export HSSB_RC_GEOM=$g; fi
    TAIL=40 run trace_$n 180 python scripts/trace_recurrent.py 512 256 64
done
unset HSSB_RC_GEOM
