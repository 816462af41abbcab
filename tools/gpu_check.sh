#!/bin/bash
# one-call GPU check: all gpu tests, then the default bench line.  Usage (under gpurun): bash tools/gpu_check.sh <tag> [bench args]
set -u
TAG=${1:-check}; shift || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/${TAG}_tests.log 2>&1; echo "gpu tests rc=$?"; grep -i "max abs err\|max |diff|" gpurun_out/${TAG}_tests.log | tail -12; tail -4 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value %.1f Mpx/s  %.1f ms/step  e2e %.1f (%s)  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('matches_device_resident'), d['clocks']))
print(d['roofline']['ms_per_step_by_kernel'], 'frac %.3f exec %.3f' % (d['roofline']['frac'], d['roofline']['frac_executed']))
print({k: round(v['frac_executed'], 3) for k, v in d['roofline']['by_epilogue'].items()})
print('facade', d.get('e2e_facade'))
print('cpu', d.get('cpu_baseline'))
PY
