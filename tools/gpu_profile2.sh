#!/bin/bash
# ncu --set full of one whole batch of the DSen2 pipeline (prep, head, 6 x (relu, res32), tail) and of the training kernels
set -u
mkdir -p gpurun_out
SMALL="python bench.py --tile 2352 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/p2_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:"conv_pair_kernel|prep_from_images" -s 15 -c 16 -o gpurun_out/r01_prof_batch $SMALL > gpurun_out/p2_ncu.log 2>&1
echo "batch capture rc=$?"
export DSEN2_TRAIN_NO_GRAPH=1
TR="python bench.py --workload train --steps 1 --warmup 1"
$TR > gpurun_out/p2_plain_train.log 2>&1 &&
ncu --set full --clock-control none -k regex:"wgrad_direct|colsum|nadam|relu_mask|mae_grad" -s 20 -c 8 -o gpurun_out/r01_prof_train $TR > gpurun_out/p2_ncu_train.log 2>&1
echo "train capture rc=$?"
