#!/bin/bash
# training-step record of one box at N GPUs: DSen2 (config 5) and VDSen2 (--deep).  Usage (under gpurun --gpus N): bash tools/gpu_multi_train.sh <tag> <N>
TAG=$1; N=$2
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29512 --workload train > gpurun_out/${TAG}_bench_train_n$N.json 2> gpurun_out/${TAG}_bench_train_n$N.err; echo "train n=$N rc=$?"
run 29514 --workload train --model vdsen2 > gpurun_out/${TAG}_bench_train_vdsen2_n$N.json 2> gpurun_out/${TAG}_bench_train_vdsen2_n$N.err; echo "train vdsen2 n=$N rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_train*_n$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], '%.1f %s' % (d['value'], d['unit']), '%.3f ms/step' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d.get('allreduce'))
    except Exception as e:
        print(f, 'ERR', e)
PY
