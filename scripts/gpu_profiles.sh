#!/bin/bash
# Round profile set: ncu launch list of the bench command + one --set full capture of every kernel of a step.
# usage: scripts/gpu_profiles.sh <tag>
set -u
OUT=gpurun_out
mkdir -p $OUT
tag=${1:-r02}
K='regex:stft_hop1|if_reassign|stats_finalize|normalise|split_planes_tiled|tc_inproj|tc_recurrent|head_kernel|metrics_kernel'
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-configs > $OUT/${tag}_launches.log 2>&1
echo "launch list rc=$?"
# the full capture profiles every kernel in its stand-alone form (HSSB_OVERLAP=0: one projection launch; ncu serialises kernels anyway).
# per step 12 matching launches, 10 of which do work (the stand-in launches of the range guard exit at once and are skipped by name / order):
# stft, reassign, stats, normalise, split, recurrent l1, [stand-in inproj + recurrent], inproj l1, recurrent l2, head, metrics
timeout -s KILL 1500 env HSSB_OVERLAP=0 ncu --set full --clock-control none --import-source on -k "$K" -s 36 -c 12 -f -o $OUT/${tag}_ncu_all \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-side-configs > $OUT/${tag}_ncu_all.log 2>&1
echo "full capture rc=$?"
