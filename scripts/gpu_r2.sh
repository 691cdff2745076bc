#!/bin/bash
# Round-2 GPU script: phases under their own hard timeouts, logs into gpurun_out/r2/.
# usage: scripts/gpu_r2.sh phase [phase...]
set -u
OUT=gpurun_out/r2
mkdir -p $OUT
run() { # name timeout cmd...
    local name=$1 to=$2; shift 2
    echo "=== $name ($(date +%T))" | tee -a $OUT/summary.txt
    timeout -s KILL $to "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "rc=$rc" | tee -a $OUT/summary.txt
    tail -n ${TAIL:-15} $OUT/$name.log | tee -a $OUT/summary.txt
    return $rc
}
rm -f $OUT/.hung
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
for p in "$@"; do
if [ -f $OUT/.hung ]; then echo "skipping $p: an earlier phase was killed by its timeout" | tee -a $OUT/summary.txt; continue; fi
case $p in
k4)      run k4 150 python -m pytest tests/test_lstm_gpu.py -q -m gpu -x -k "k4" || { [ $? -eq 137 ] && touch $OUT/.hung; } ;;
quick)   run quick 300 python -m pytest tests/test_lstm_gpu.py -q -m gpu -x -k "golden or odd_batch or fused_and or input_range" || { [ $? -eq 137 ] && touch $OUT/.hung; } ;;
tests)   run tests 1500 python -m pytest tests -q -m gpu -x -rs --durations=8 || { [ $? -eq 137 ] && touch $OUT/.hung; } ;;
newtests) run newtests 900 python -m pytest tests -q -m gpu -x -s -k "config1 or config3_full or config4_shard_full or input_range or weights_outside or copies or metric_state or frames_of or recording_to or second_device" ;;
lstm)    run lstm 900 python -m pytest tests/test_lstm_gpu.py -q -m gpu -x ;;
fsst)    run fsst 600 python -m pytest tests/test_fsst_gpu.py -q -m gpu -x ;;
smoke)   run smoke 300 python -c "import __graft_entry__ as g; g.smoke()" ;;
bench)   run bench 900 python bench.py ; cp $OUT/bench.log $OUT/bench_$(date +%H%M%S).json ;;
benchq)  run benchq 600 python bench.py --no-cpu-baseline --no-side-configs ;;
ref)     run ref 600 python bench.py --impl reference --steps 2 --warmup 1 ;;
launches) run launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-side-configs ;;
ncu)     run ncu 1200 ncu --set full --clock-control none --import-source on -s 60 -c 24 -o $OUT/ncu_all -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-side-configs ;;
race)    run race 1500 bash scripts/sanitize_r2.sh ;;
trace)   run trace 300 python scripts/trace_recurrent.py 512 256 64 ;;
*)       echo "unknown phase $p" ;;
esac
done
