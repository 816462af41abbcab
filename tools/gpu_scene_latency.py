"""600x600 scene (BASELINE configs 1 / 2) through the public facade, numpy in -> numpy out, and through bench.py."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsen2_b200 import supres
from dsen2_b200.DSen2Net import s2model

rng = np.random.RandomState(0)
d10 = (rng.rand(600, 600, 4) * 4000).astype(np.float32)
d20 = (rng.rand(300, 300, 6) * 4000).astype(np.float32)
d60 = (rng.rand(100, 100, 2) * 4000).astype(np.float32)
m20 = s2model(((4, None, None), (6, None, None)), 6, 128, seed=0)
m60 = s2model(((4, None, None), (6, None, None), (2, None, None)), 6, 128, seed=0)
for name, fn in (("DSen2_20", lambda: supres.DSen2_20(d10, d20, model=m20)),
                 ("DSen2_60", lambda: supres.DSen2_60(d10, d20, d60, model=m60))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t) * 100
    print("%s numpy in -> numpy out, 600x600 scene: %.2f ms = %.1f output Mpixel/s" % (name, ms, 0.36 / ms * 1e3), flush=True)
