#!/bin/bash
# A/B of the trunk formats on the same box: q8 (default) vs fp32, tile 2352 and the full tile
mkdir -p gpurun_out
for fmt in q8 fp32 q8 fp32; do
  DSEN2_TRUNK=$fmt timeout 300 python bench.py --tile 2352 --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$fmt tile2352 %.1f Mpx/s %.2f ms e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['roofline']['ms_per_step_by_kernel'], 'chk', d['checksum'])"
done
for fmt in q8 fp32; do
  DSEN2_TRUNK=$fmt timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$fmt full %.1f Mpx/s %.2f ms e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['roofline']['ms_per_step_by_kernel'], 'chk', d['checksum'])"
done
