#!/bin/bash
# `ncu --set full` of the training step's kernels (DSen2, config 5; eager launches so that ncu sees them).
# Usage (under gpurun, one GPU): bash tools/gpu_profile_train_full.sh <tag>
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --workload train --steps 1 --warmup 2"
export DSEN2_TRAIN_NO_GRAPH=1
$CMD > gpurun_out/${TAG}_trainfull_plain.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/${TAG}_trainfull_plain.log; exit 1; }
# 30 consecutive convolution launches of the third step on: first layer, 6 x (RELU, RESIDUAL32), last layer, then the backward
# pass (RESIDUAL32 with scale 1, MASK)
ncu --set full --clock-control none -k regex:conv_pair_kernel -s 64 -c 30 -o gpurun_out/${TAG}_train_conv_prof $CMD > gpurun_out/${TAG}_train_conv_ncu.log 2>&1
echo "conv capture rc=$?"
ncu --set full --clock-control none -k regex:"wgrad_direct_kernel|nadam|pack_trunk_layers_kernel|mae_grad_kernel" -s 32 -c 8 \
  -o gpurun_out/${TAG}_train_other_prof $CMD > gpurun_out/${TAG}_train_other_ncu.log 2>&1
echo "wgrad / nadam capture rc=$?"
