#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/gpu_diag.py --one model_timing > gpurun_out/plain_model_timing.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_pair_kernel -s 30 -c 4 -o gpurun_out/r01_prof_pair3 python tools/gpu_diag.py --one model_timing > gpurun_out/ncu_pair3.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_pair3.log
