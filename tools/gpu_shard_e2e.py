"""One rank's share of a full tile (rank 3 of 8) on ONE GPU: device-resident time vs the host pipeline (pinned uint16 in,
float32 out).  Separates the pipeline's own overhead from PCIe / host-memory contention between ranks of a real 8-GPU run."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsen2_b200 import sharding, supres
from dsen2_b200.DSen2Net import s2model

T = 10980
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(1)
d10 = torch.randint(0, 12000, (T, T, 4), generator=g, device=dev).to(torch.float32)
d20 = torch.randint(0, 12000, (T // 2, T // 2, 6), generator=g, device=dev).to(torch.float32)
model = s2model(((4, None, None), (6, None, None)), 6, 128, seed=0)
first, count = sharding.shard_range(9801, 3, 8)
out = torch.zeros((T, T, 6), device=dev)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms_dev = timed(lambda: supres.super_resolve_device(model, d10, d20, first_patch=first, num_patches=count, out=out))
h10 = torch.empty((T, T, 4), dtype=torch.uint16).pin_memory()
h20 = torch.empty((T // 2, T // 2, 6), dtype=torch.uint16).pin_memory()
h10.numpy()[...] = d10.cpu().numpy()
h20.numpy()[...] = d20.cpu().numpy()
hout = torch.zeros((T, T, 6)).pin_memory()
del d10, d20
pipe = supres.HostPipeline(model, T, T, dtype=torch.uint16)
ms_e2e = timed(lambda: pipe.run(h10, h20, hout=hout, first_patch=first, num_patches=count))
print('rank 3 of 8 on one GPU: device-resident %.2f ms, host pipeline %.2f ms (+%.2f), %d chunks, H2D %.1f MB, D2H %.1f MB'
      % (ms_dev, ms_e2e, ms_e2e - ms_dev, len(pipe._plan(first, count)), pipe.h2d_bytes / 1e6, pipe.d2h_bytes / 1e6))
for label, nbytes in (('D2H', 362e6), ('H2D', 180e6)):
    a = torch.empty(int(nbytes) // 4, device=dev)
    b = torch.empty(int(nbytes) // 4).pin_memory()
    src, dst = (a, b) if label == 'D2H' else (b, a)
    ms = timed(lambda: dst.copy_(src, non_blocking=True))
    print('%s of %.0f MB pinned: %.2f ms = %.1f GB/s' % (label, nbytes / 1e6, ms, nbytes / ms / 1e6))
