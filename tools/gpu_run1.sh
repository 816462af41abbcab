#!/bin/bash
# first bring-up of the CTA-pair kernels: tests in separate processes with timeouts, then timings
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py -x -q -m gpu > gpurun_out/t_pair.log 2>&1; echo "pair tests rc=$?"; tail -15 gpurun_out/t_pair.log
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_patches.py -x -q -m gpu > gpurun_out/t_rest.log 2>&1; echo "rest rc=$?"; tail -8 gpurun_out/t_rest.log
timeout 300 python tools/gpu_diag.py pair_timing model_timing > gpurun_out/diag2.log 2>&1; cat gpurun_out/diag2.log
