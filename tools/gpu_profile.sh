#!/bin/bash
# ncu evidence for the current build (run under gpurun, one GPU).  Usage: tools/gpu_profile.sh <tag>
# 1. launch list (gpu__time_duration per launch) of a small-tile bench run; 2. --set full capture of the trunk kernels.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
SMALL="python bench.py --tile 2352 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/${TAG}_plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$SMALL > gpurun_out/${TAG}_plain_small2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 20 -c 6 -o gpurun_out/${TAG}_prof_conv $SMALL > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
