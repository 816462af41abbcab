#!/bin/bash
# ncu --set full of the head (first conv_pair launch of a batch) and the tail inside bench.py --tile 2352
set -u
TAG=${1:-r01h16}
mkdir -p gpurun_out
SMALL="python bench.py --tile 2352 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 27 -c 2 -o gpurun_out/${TAG}_prof $SMALL > gpurun_out/${TAG}_ncu.log 2>&1
echo "capture rc=$?"
