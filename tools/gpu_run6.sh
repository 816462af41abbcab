#!/bin/bash
set -u
mkdir -p gpurun_out
for b in 12 16 24 32 48 64 96 128; do
timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --batch $b > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_b$b.json').read().strip().splitlines()[-1])
k=d['roofline']['ms_per_step_by_kernel']
print('batch $b value %.1f ms %.1f e2e %.1f sm %s | '%(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks']['sm_mhz'])+' '.join('%s=%.0f'%(a,b) for a,b in k.items()))
PY
done
