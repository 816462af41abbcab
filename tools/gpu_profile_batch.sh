#!/bin/bash
# ncu --set full of ONE WHOLE BATCH of the current DSen2 pipeline (prep16, head16, 6 x (RELU, RESIDUALQ), tail) inside bench.py --tile 2352
set -u
TAG=${1:-r01batch}
mkdir -p gpurun_out
SMALL="python bench.py --tile 2352 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:"conv_pair_kernel|conv_tail_swapped|prep16_from_images" -s 15 -c 15 -o gpurun_out/${TAG}_prof $SMALL > gpurun_out/${TAG}_ncu.log 2>&1
echo "batch capture rc=$?"
