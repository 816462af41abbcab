#!/bin/bash
# one-call GPU check: all gpu tests, then the default bench line.  Usage (under gpurun): bash tools/gpu_check.sh <tag>
set -u
TAG=${1:-check}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value %.1f Mpx/s  %.1f ms/step  e2e %.1f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']))
print(d['roofline']['ms_per_step_by_kernel'], 'frac %.3f exec %.3f' % (d['roofline']['frac'], d['roofline']['frac_executed']))
PY
