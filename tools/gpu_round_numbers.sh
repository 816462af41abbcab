#!/bin/bash
# the other single-GPU configurations + the measured error against the oracle, for DESIGN.md / profiles
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -q -s -k "oracle" 2>&1 | grep -i "max abs err\|passed\|failed" | tee gpurun_out/r02r_err.txt
timeout 600 python bench.py --path 60 > gpurun_out/r02r_bench_p60.json 2> gpurun_out/r02r_bench_p60.err; echo "p60 rc=$?"
timeout 900 python bench.py --model vdsen2 --tile 3360 --steps 2 --warmup 3 > gpurun_out/r02r_bench_vd.json 2> gpurun_out/r02r_bench_vd.err; echo "vd rc=$?"
timeout 600 python bench.py --workload train > gpurun_out/r02r_bench_train.json 2> gpurun_out/r02r_bench_train.err; echo "train rc=$?"
python - <<PY
import json
for f in ('p60','vd','train'):
    try:
        d=json.loads(open('gpurun_out/r02r_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['metric'], '%.1f'%d['value'], d['unit'], '%.2f ms'%d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
