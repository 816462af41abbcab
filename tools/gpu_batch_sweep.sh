#!/bin/bash
# device-batch sweep: does a working set that fits the 126 MB L2 beat the launch overhead of small batches?
mkdir -p gpurun_out
for b in 2 3 4 5 6 8 12 16 21 42 84; do
  timeout 120 python bench.py --tile 2352 --batch $b --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('batch %3d  %.1f Mpx/s  %.2f ms  e2e %.1f' % ($b, d['value'], d['ms_per_step'], d['e2e']['value']))" \
    | tee -a gpurun_out/batch_sweep.txt
done
