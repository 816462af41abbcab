#!/usr/bin/env python
"""Where does the data-parallel VDSen2 step spend its time?  (torchrun, N >= 2)  CUDA events around the pieces of the
overlapped schedule: gradient graph A, [all-reduce of the late bucket || gradient graph B], all-reduce of the early bucket,
update graph -- and the same pieces run one after the other."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsen2_b200.DSen2Net import s2model          # noqa: E402
from dsen2_b200.train import Nadam, Trainer, allreduce_gradients      # noqa: E402

rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
deep = '--dsen2' not in sys.argv
L, F, n = (32, 256, 8) if deep else (6, 128, 128)
model = s2model(((4, None, None), (6, None, None)), num_layers=L, feature_size=F, seed=0)
tr = Trainer(model, Nadam(lr=1e-4), device=dev)
if not tr._overlap_allreduce(n, 32):
    tr._overlap_allreduce = lambda n_, P_: True      # force the two-part schedule to look at it
g = torch.Generator().manual_seed(rank)
xs = [torch.rand((n, c, 32, 32), generator=g).mul_(2.5).to(dev) for c in (4, 6)]
y = torch.rand((n, 6, 32, 32), generator=g).mul_(2.5).to(dev)
for _ in range(4):
    tr.train_step(xs, y)
torch.cuda.synchronize()
ga, gb, g2 = tr._graphs[(n, 32, 2)]
off = int(tr.offsets[2 * (1 + 2 * tr.split_l)])
ev = lambda: torch.cuda.Event(enable_timing=True)


def run(overlap, iters=20):
    acc = [0.0] * 5
    for _ in range(iters):
        e = [ev() for _ in range(6)]
        dist.barrier()
        torch.cuda.synchronize()
        e[0].record(); ga.replay(); e[1].record()
        if overlap:
            late = allreduce_gradients(tr.grads[off:], None, async_op=True)
            gb.replay(); e[2].record()
            late.wait(); e[3].record()
        else:
            gb.replay(); e[2].record()
            allreduce_gradients(tr.grads[off:]); e[3].record()
        allreduce_gradients(tr.grads[:off]); e[4].record()
        g2.replay(); e[5].record()
        torch.cuda.synchronize()
        for i in range(5):
            acc[i] += e[i].elapsed_time(e[i + 1]) * 1e3 / iters
    return acc


for mode in (True, False, True, False):
    a = run(mode)
    if rank == 0:
        print('%-9s graph A %7.1f us | graph B %7.1f | %s %7.1f | early all-reduce %6.1f | update %6.1f | total %7.1f' % (
            'overlap' if mode else 'serial', a[0], a[1], 'wait for late  ' if mode else 'late all-reduce', a[2], a[3], a[4], sum(a)), flush=True)
dist.destroy_process_group()
