"""GPU bring-up diagnostics (development tool, not a test): each check runs in its own process with a timeout
so that a hung kernel cannot take the whole gpurun call with it.  Usage: python tools/gpu_diag.py [check ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _time(torch, fn, iters=10, warm=3):
    if os.environ.get('DSEN2_DIAG_ONCE'):          # under ncu: one launch of everything
        iters, warm = 1, 0
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def check_layers():
    """One launch of every layer kind of the DSen2 / VDSen2 forward on 84 patches of 128 x 128: ms and executed TFLOP/s."""
    import numpy as np
    import torch
    from dsen2_b200.DSen2Net import s2model
    rng = np.random.RandomState(0)
    for F, L, n in ((128, 6, 84), (256, 2, 84)):
        model = s2model(((4, None, None), (6, None, None)), num_layers=L, feature_size=F, seed=0)
        xs = [torch.from_numpy(rng.rand(n, c, 128, 128).astype(np.float32)).cuda() for c in (4, 6)]
        for _ in range(3):
            model.forward_device(xs)
        timers = {}
        model.forward_device(xs, timers=timers)
        torch.cuda.synchronize()
        for kind, ev in timers.items():
            ms = float(np.mean([a.elapsed_time(b) for a, b, _ in ev]))
            fl = 2.0 * 9 * F * F * 128 * 128 * n if kind.startswith('conv_res') else 0.0
            print('F=%d %-10s %.3f ms%s' % (F, kind, ms, '  %.0f TFLOP/s executed' % (fl / ms / 1e9) if fl else ''), flush=True)


def check_hbm_kernels():
    """Standalone drop-in kernels behind utils.patches / utils.imresize at full-tile sizes: ms and ALGORITHMIC GB/s
    (extract reads + writes the patch stack, bilinear reads p^2 + writes (2p)^2 per plane, recompose
    reads + writes the owned pixels, bicubic reads the input + writes the float64 output)."""
    import torch
    from dsen2_b200 import imresize, patches
    T = 10980
    gen = torch.Generator(device='cuda').manual_seed(1)
    d10 = torch.rand((T, T, 4), device='cuda', generator=gen) * 4000
    d20 = torch.rand((T // 2, T // 2, 6), device='cuda', generator=gen) * 4000
    n = 99 * 20                                             # 20 patch rows of the 20 m path
    out10 = torch.empty((n, 4, 128, 128), device='cuda')
    ms = _time(torch, lambda: patches.extract_patches_device(d10, 2, 64, 4, 0, n, 2000.0, out10))
    print('extract 10 m (n=%d, 4x128x128, /2000) : %.3f ms  %.0f GB/s' % (n, ms, 2 * out10.numel() * 4 / ms / 1e6), flush=True)
    out20 = torch.empty((n, 6, 64, 64), device='cuda')
    ms = _time(torch, lambda: patches.extract_patches_device(d20, 1, 64, 4, 0, n, 1.0, out20))
    print('extract 20 m (n=%d, 6x64x64)          : %.3f ms  %.0f GB/s' % (n, ms, 2 * out20.numel() * 4 / ms / 1e6), flush=True)
    up = torch.empty((n, 6, 128, 128), device='cuda')
    ms = _time(torch, lambda: patches.bilinear_up_device(out20, 2, 2000.0, up))
    print('bilinear mirror x2 (%d planes 64->128) : %.3f ms  %.0f GB/s' % (n * 6, ms, (out20.numel() + up.numel()) * 4 / ms / 1e6), flush=True)
    canvas = torch.zeros((T, T, 6), device='cuda')
    ms = _time(torch, lambda: patches.recompose_device(up, 8, T, T, 0, 2000.0, canvas))
    owned = n * 112 * 112 * 6 * 4
    print('recompose (n=%d, 6 bands, x2000)      : %.3f ms  %.0f GB/s' % (n, ms, 2 * owned / ms / 1e6), flush=True)
    del out10, up, canvas, d10
    torch.cuda.empty_cache()
    ms = _time(torch, lambda: imresize.imresize_device(d20, (2.0, 2.0), (T, T)), iters=4, warm=2)
    print('bicubic imresize x2 (5490^2x6 -> f64)  : %.3f ms  %.0f GB/s  (incl. tap tables on the host)' % (
        ms, (d20.numel() * 4 + T * T * 6 * 8) / ms / 1e6), flush=True)
    d60 = torch.rand((T // 6, T // 6, 2), device='cuda', generator=gen) * 4000
    ms = _time(torch, lambda: imresize.imresize_device(d60, (6.0, 6.0), (T, T)), iters=4, warm=2)
    print('bicubic imresize x6 (1830^2x2 -> f64)  : %.3f ms  %.0f GB/s  (incl. tap tables on the host)' % (
        ms, (d60.numel() * 4 + T * T * 2 * 8) / ms / 1e6), flush=True)


CHECKS = {k[6:]: v for k, v in globals().items() if k.startswith('check_')}

if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--one':
        CHECKS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(CHECKS)
    for nme in names:
        t = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--one', nme], timeout=240)
            print('[diag] %s exit %d in %.1fs' % (nme, r.returncode, time.time() - t), flush=True)
        except subprocess.TimeoutExpired:
            print('[diag] %s TIMEOUT' % nme, flush=True)
