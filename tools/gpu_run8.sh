#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"; tail -6 gpurun_out/t_all.log
timeout 600 python bench.py > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_v4.err; cat gpurun_out/bench_v4.json
