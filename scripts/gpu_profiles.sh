#!/bin/bash
# Round profile set: ncu launch list of the bench command + one --set full capture of every kernel of a step.
# usage: scripts/gpu_profiles.sh <tag>
set -u
OUT=gpurun_out
mkdir -p $OUT
tag=${1:-r01c}
K='regex:stft_hop1|if_reassign|stats_finalize|normalise|split_planes|tc_inproj|tc_recurrent|head_kernel|confusion_kernel'
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${tag}_launches.log 2>&1
echo "launch list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "$K" -s 30 -c 10 -f -o $OUT/${tag}_ncu_all \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${tag}_ncu_all.log 2>&1
echo "full capture rc=$?"
