#!/bin/bash
# multi-GPU record of one box: DSen2 full tile (config 3), optionally VDSen2 full tile (config 4), training step (config 5).
# Usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> [vdsen2]
TAG=$1; N=$2; VD=${3:-}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29511 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "dsen2 n=$N rc=$?"
run 29512 --workload train > gpurun_out/${TAG}_bench_train_n$N.json 2> gpurun_out/${TAG}_bench_train_n$N.err; echo "train n=$N rc=$?"
if [ -n "$VD" ]; then
  run 29513 --model vdsen2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_vdsen2_n$N.json 2> gpurun_out/${TAG}_bench_vdsen2_n$N.err; echo "vdsen2 n=$N rc=$?"
  run 29514 --workload train --model vdsen2 > gpurun_out/${TAG}_bench_train_vdsen2_n$N.json 2> gpurun_out/${TAG}_bench_train_vdsen2_n$N.err; echo "train vdsen2 n=$N rc=$?"
fi
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*_n$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['metric'], '%.1f %s' % (d['value'], d['unit']), '%.2f ms/step' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d.get('allreduce'), (d.get('clocks') or {}).get('sm_mhz'))
    except Exception as e:
        print(f, 'ERR', e)
PY
