#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/t_pair.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_pair.log
timeout 300 python tools/gpu_diag.py pair_debug_sweep pair_timing model_timing > gpurun_out/diag3.log 2>&1; cat gpurun_out/diag3.log
