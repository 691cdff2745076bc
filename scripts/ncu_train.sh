#!/bin/bash
# ncu --set full capture of the training step's tensor-core kernels: K4 (layer 1, layer 2), K5m TRAIN x 2, K5b x 2 of one step.
# usage: scripts/ncu_train.sh <tag>
set -u
OUT=gpurun_out
mkdir -p $OUT
tag=${1:-r02}
K='regex:tc_bptt|tc_recurrent_mc|tc_inproj|split_slots|split_tf32|ce_head|clip_adam'
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k "$K" -c 22 -f -o $OUT/${tag}_ncu_train \
    python scripts/time_train.py > $OUT/${tag}_ncu_train.log 2>&1
echo "training capture rc=$?"
