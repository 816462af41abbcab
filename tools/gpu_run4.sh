#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/t_pair.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_pair.log
timeout 300 python tools/gpu_diag.py pair_debug_sweep pair_timing model_timing > gpurun_out/diag4.log 2>&1; cat gpurun_out/diag4.log
python tools/gpu_diag.py --one pair_relu_only > gpurun_out/plain_relu_only.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_pair_kernel -s 10 -c 6 -o gpurun_out/r01_prof_pair2 python tools/gpu_diag.py --one pair_relu_only > gpurun_out/ncu_pair2.log 2>&1
echo "ncu rc=$?"
