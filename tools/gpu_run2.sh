#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py -x -q -m gpu > gpurun_out/t_pair.log 2>&1; echo "pair tests rc=$?"; tail -5 gpurun_out/t_pair.log
python tools/gpu_diag.py --one pair_timing > gpurun_out/plain_pair_timing.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_pair_kernel -s 10 -c 6 -o gpurun_out/r01_prof_pair python tools/gpu_diag.py --one pair_timing > gpurun_out/ncu_pair.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/ncu_pair.log
