#!/bin/bash
# A/B of an experiment env var on one box: alternating runs of the full-tile bench.  Usage: bash tools/gpu_exp.sh <tag> VAR=VALUE
TAG=$1; KV=$2
mkdir -p gpurun_out
for i in 1 2; do
  for mode in base exp; do
    if [ $mode = exp ]; then export $KV; else unset ${KV%%=*}; fi
    timeout 600 python bench.py --steps 4 --warmup 2 --no-facade --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${TAG}_${mode}_$i.json
    python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_${mode}_$i.json').read())
print('$mode $i: %.1f ms/step' % d['ms_per_step'], {k: round(v,1) for k,v in d['roofline']['ms_per_step_by_kernel'].items()}, d['clocks']['sm_mhz'])
PY
  done
done
