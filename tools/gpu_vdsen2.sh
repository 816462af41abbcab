#!/bin/bash
# VDSen2 (config 4) on one GPU: a 3360^2 tile (900 patches) by default.  Usage (under gpurun): bash tools/gpu_vdsen2.sh <tag> [tile]
TAG=${1:-vd}; TILE=${2:-3360}
mkdir -p gpurun_out
timeout 1200 python bench.py --model vdsen2 --tile $TILE --steps 2 --warmup 2 --no-facade --no-cpu-baseline > gpurun_out/${TAG}_bench_vd.json 2> gpurun_out/${TAG}_bench_vd.err; echo "vd rc=$?"; tail -3 gpurun_out/${TAG}_bench_vd.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_vd.json').read().strip().splitlines()[-1])
print('value %.2f Mpx/s  %.1f ms/step  e2e %.2f (%s)  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('matches_device_resident'), d['clocks']))
print(d['roofline']['ms_per_step_by_kernel'], 'frac %.3f exec %.3f' % (d['roofline']['frac'], d['roofline']['frac_executed']))
print({k: round(v['frac_executed'], 3) for k, v in d['roofline']['by_epilogue'].items()}, 'whole step TF', d['tflops_executed_whole_step'])
PY
