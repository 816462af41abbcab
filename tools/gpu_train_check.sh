#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/t_train.log 2>&1; echo "train tests rc=$?"; tail -25 gpurun_out/t_train.log
