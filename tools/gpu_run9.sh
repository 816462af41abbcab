#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/t_pair.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_pair.log
timeout 300 python tools/gpu_diag.py pair_timing > gpurun_out/diag9.log 2>&1; grep "n=64" gpurun_out/diag9.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_v5.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v5.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['clocks'])
print(d['roofline']['ms_per_step_by_kernel'], d['roofline']['frac'], d['roofline']['frac_executed'])
PY
