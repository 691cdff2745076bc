#!/bin/bash
# First-contact GPU script: every phase under its own hard timeout, logs into gpurun_out/.
# usage: scripts/gpu_check.sh [phases...]   (default: all)
set -u
OUT=gpurun_out
mkdir -p $OUT
PH="${@:-fsst simt k4 tc bench_simt bench launches}"
run() { # name timeout cmd...
    local name=$1 to=$2; shift 2
    echo "=== $name ($(date +%T))" | tee -a $OUT/summary.txt
    timeout -s KILL $to "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "rc=$rc" | tee -a $OUT/summary.txt
    tail -n 12 $OUT/$name.log | tee -a $OUT/summary.txt
    return $rc
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
for p in $PH; do
case $p in
fsst)  run fsst 600 python -m pytest tests/test_fsst_gpu.py -q -m gpu -x ;;
simt)  run simt 600 python -m pytest tests/test_lstm_gpu.py -q -m gpu -k "simt or confusion" ;;
k4)    run k4 180 python -m pytest tests/test_lstm_gpu.py -q -m gpu -k "k4" ;;
tc)    run tc 300 python -m pytest tests/test_lstm_gpu.py -q -m gpu -k "not simt and not k4 and not confusion" ;;
bench_simt) HSSB_LSTM_IMPL=simt run bench_simt 400 python bench.py --steps 2 --warmup 1 --windows 128 --no-cpu-baseline ;;
bench) run bench 600 python bench.py ;;
launches) run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline ;;
smoke) run smoke 300 python -c "import __graft_entry__ as g; g.smoke()" ;;
all)   run all 900 python -m pytest tests -q -m gpu -x ;;
esac
done
