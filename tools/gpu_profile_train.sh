#!/bin/bash
# ncu launch lists of the training step (run under gpurun, one GPU).  Usage: tools/gpu_profile_train.sh <tag>
# DSEN2_TRAIN_NO_GRAPH=1 makes the trainer launch eagerly so that ncu sees the individual kernels.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
for M in dsen2 vdsen2; do
  CMD="python bench.py --workload train --model $M --steps 2 --warmup 2"
  DSEN2_TRAIN_NO_GRAPH=1 $CMD > gpurun_out/${TAG}_train_${M}_plain.log 2>&1 &&
  DSEN2_TRAIN_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/${TAG}_train_${M}_launches.csv $CMD > gpurun_out/${TAG}_train_${M}_ncu.log 2>&1
  echo "$M launch list rc=$?"
  python tools/ncu_summary.py list gpurun_out/${TAG}_train_${M}_launches.csv > gpurun_out/${TAG}_train_${M}_launch_list.txt 2>&1
  head -30 gpurun_out/${TAG}_train_${M}_launch_list.txt
done
