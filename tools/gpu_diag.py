"""GPU bring-up diagnostics (development tool, not a test): each check runs in its own process with a timeout
so that a hung kernel cannot take the whole gpurun call with it.  Usage: python tools/gpu_diag.py [check ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def check_rowshift():
    import numpy as np
    import torch
    from dsen2_b200 import _capi
    lib = _capi.lib()
    rng = np.random.RandomState(0)
    a = (rng.rand(160, 64).astype(np.float32) - 0.5).astype(np.float16)
    b = (rng.rand(128, 64).astype(np.float32) - 0.5).astype(np.float16)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    for mode in (0, 1):
        res = []
        for shift in list(range(0, 18)) + [24, 31, 32]:
            out = torch.zeros((128, 128), device='cuda')
            _capi.check(lib.dsen2_debug_umma_rowshift(_capi.ptr(ta), 160, _capi.ptr(tb), shift, mode, _capi.ptr(out),
                                                      _capi.stream_ptr()), 'umma')
            torch.cuda.synchronize()
            ref = a[shift:shift + 128].astype(np.float32) @ b.astype(np.float32).T
            err = float(np.abs(out.cpu().numpy() - ref).max())
            res.append((shift, round(err, 5)))
        print('base_offset_mode', mode, res, flush=True)


def check_conv_timing():
    import numpy as np
    import torch
    from dsen2_b200 import _capi
    lib = _capi.lib()
    for F, n in ((128, 64), (256, 32)):
        P = 128
        x = torch.randn((n, P, P, F), device='cuda').half()
        w = (torch.rand((9, F, F), device='cuda') - 0.5).half()
        b = torch.zeros(F, device='cuda')
        hi = torch.empty_like(x)
        def run():
            _capi.check(lib.dsen2_conv3x3(_capi.ptr(x), _capi.ptr(w), _capi.ptr(b), n, P, P, F, F, 9, 0, None, None,
                                          0.0, _capi.ptr(hi), None, None, None, 0, _capi.stream_ptr()), 'conv')
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 10
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * n * P * P * 9 * F * F
        print('conv3x3 F=%d n=%d: %.3f ms  %.1f TFLOP/s' % (F, n, ms, fl / ms / 1e9), flush=True)


def check_model_timing():
    import numpy as np
    import torch
    from dsen2_b200.DSen2Net import s2model
    m = s2model(((4, None, None), (6, None, None)), 6, 128, seed=0)
    n, P = 64, 128
    xs = [torch.rand((n, 4, P, P), device='cuda'), torch.rand((n, 6, P, P), device='cuda')]
    for _ in range(2):
        m.forward_device(xs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m.forward_device(xs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 3575808.0 * n * P * P
    print('DSen2-20 forward n=%d: %.3f ms  %.1f TFLOP/s  %.2f Mpx/s(patch px)' % (n, ms, fl / ms / 1e9, n * P * P / ms / 1e3),
          flush=True)


def _time(torch, fn, iters=10, warm=3):
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


def check_pair_timing():
    """Per-layer timings of the DSen2 fast path (CTA-pair kernels) next to the single-CTA kernel."""
    import torch
    from dsen2_b200 import _capi
    lib, ptr = _capi.lib(), _capi.ptr
    st = _capi.stream_ptr()
    F, P = 128, 128
    for n in [int(v) for v in os.environ.get('DSEN2_DIAG_N', '64,8').split(',')]:
        x = torch.randn((n, P, P, F), device='cuda').half()
        w = (torch.rand((9, F, F), device='cuda') - 0.5).half()
        b = torch.zeros(F, device='cuda')
        hi, lo, t = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        fl = 2.0 * n * P * P * 9 * F * F
        for v1 in (0, 1):
            lib.dsen2_debug_force_v1(v1)
            ms = _time(torch, lambda: _capi.check(lib.dsen2_conv3x3(ptr(x), ptr(w), ptr(b), n, P, P, F, F, 9, 0, None, None,
                                                                    0.0, ptr(t), None, None, None, 0, st), 'relu'))
            print('n=%d %s RELU     : %.3f ms  %.1f TFLOP/s' % (n, 'v1  ' if v1 else 'pair', ms, fl / ms / 1e9), flush=True)
            ms = _time(torch, lambda: _capi.check(lib.dsen2_conv3x3(ptr(t), ptr(w), ptr(b), n, P, P, F, F, 9, 1, ptr(hi),
                                                                    ptr(lo), 0.1, ptr(hi), ptr(lo), None, None, 0, st), 'res'))
            print('n=%d %s RESIDUAL : %.3f ms  %.1f TFLOP/s' % (n, 'v1  ' if v1 else 'pair', ms, fl / ms / 1e9), flush=True)
        lib.dsen2_debug_force_v1(0)
        xin_hi = torch.randn((n, P, P, 64), device='cuda').half()
        xin_lo = (torch.randn((n, P, P, 64), device='cuda') * 1e-3).half()
        wh = (torch.rand((3, 2 * F, 64), device='cuda') - 0.5).half()
        x32 = torch.zeros((n, P, P // 8, F // 4, 8, 4), device='cuda')
        ms = _time(torch, lambda: _capi.check(lib.dsen2_conv_res32(ptr(t), ptr(w), ptr(b), n, P, P, 0.1, ptr(x32), ptr(hi),
                                                                   None, st), 'res32'))
        print('n=%d pair RESIDUAL32 (fp32 trunk): %.3f ms  %.1f TFLOP/s' % (n, ms, fl / ms / 1e9), flush=True)
        ms = _time(torch, lambda: _capi.check(lib.dsen2_conv_head(ptr(xin_hi), ptr(xin_lo), ptr(wh), ptr(b), n, P, P, F,
                                                                  ptr(hi), None, ptr(x32), st), 'head'))
        print('n=%d head (split)      : %.3f ms' % (n, ms), flush=True)
        wt = (torch.rand((9, 32, F), device='cuda') - 0.5).half()
        pred = torch.empty((n, 6, P, P), device='cuda')
        ms = _time(torch, lambda: _capi.check(lib.dsen2_conv_tail(ptr(hi), ptr(lo), ptr(wt), ptr(b), ptr(xin_hi),
                                                                  ptr(xin_lo), 4, 6, n, P, P, ptr(pred), st), 'tail'))
        print('n=%d tail (split)      : %.3f ms' % (n, ms), flush=True)
        H = 112 * max(8, int(n ** 0.5) + 1)
        d10 = torch.rand((H, H, 4), device='cuda') * 4000
        d20 = torch.rand((H // 2, H // 2, 6), device='cuda') * 4000
        ms = _time(torch, lambda: _capi.check(lib.dsen2_prep_from_images(ptr(d10), ptr(d20), None, H, H, 128, 8, 0, n,
                                                                         2000.0, ptr(xin_hi), ptr(xin_lo), st), 'prep'))
        print('n=%d prep_from_images  : %.3f ms  (%.0f GB/s written)' % (n, ms, n * P * P * 256 / ms / 1e6), flush=True)


def check_pair_relu_only():
    """RELU / RESIDUAL trunk layer timing only (used with DSEN2_PAIR_DEBUG to bound the MMA rate)."""
    import torch
    from dsen2_b200 import _capi
    lib, ptr = _capi.lib(), _capi.ptr
    st = _capi.stream_ptr()
    F, P, n = 128, 128, 64
    x = torch.randn((n, P, P, F), device='cuda').half()
    w = (torch.rand((9, F, F), device='cuda') - 0.5).half()
    b = torch.zeros(F, device='cuda')
    hi, lo, t = torch.zeros_like(x), torch.zeros_like(x), torch.empty_like(x)
    fl = 2.0 * n * P * P * 9 * F * F
    ms = _time(torch, lambda: _capi.check(lib.dsen2_conv3x3(ptr(x), ptr(w), ptr(b), n, P, P, F, F, 9, 0, None, None,
                                                            0.0, ptr(t), None, None, None, 0, st), 'relu'))
    print('DEBUG=%s RELU     : %.3f ms  %.1f TFLOP/s' % (os.environ.get('DSEN2_PAIR_DEBUG', '0'), ms, fl / ms / 1e9), flush=True)
    ms = _time(torch, lambda: _capi.check(lib.dsen2_conv3x3(ptr(t), ptr(w), ptr(b), n, P, P, F, F, 9, 1, ptr(hi),
                                                            ptr(lo), 0.1, ptr(hi), ptr(lo), None, None, 0, st), 'res'))
    print('DEBUG=%s RESIDUAL : %.3f ms  %.1f TFLOP/s' % (os.environ.get('DSEN2_PAIR_DEBUG', '0'), ms, fl / ms / 1e9), flush=True)


def check_pair_res32_only():
    import torch
    from dsen2_b200 import _capi
    lib, ptr = _capi.lib(), _capi.ptr
    st = _capi.stream_ptr()
    F, P, n = 128, 128, 64
    x = torch.randn((n, P, P, F), device='cuda').half()
    w = (torch.rand((9, F, F), device='cuda') - 0.5).half()
    b = torch.zeros(F, device='cuda')
    hi, t = torch.zeros_like(x), torch.empty_like(x)
    x32 = torch.zeros((n, P, P // 8, F // 4, 8, 4), device='cuda')
    fl = 2.0 * n * P * P * 9 * F * F
    def both():
        _capi.check(lib.dsen2_conv3x3(ptr(hi), ptr(w), ptr(b), n, P, P, F, F, 9, 0, None, None, 0.0, ptr(t), None, None, None, 0, st), 'relu')
        _capi.check(lib.dsen2_conv_res32(ptr(t), ptr(w), ptr(b), n, P, P, 0.1, ptr(x32), ptr(hi), None, st), 'res32')
    ms = _time(torch, both, iters=12)
    print('PF_A=%s PF_X=%s DEFER=%s  resblock (RELU + RESIDUAL32): %.3f ms  %.1f TFLOP/s' % (
        os.environ.get('DSEN2_PAIR_PF_A', 'd'), os.environ.get('DSEN2_PAIR_PF_X', 'd'), os.environ.get('DSEN2_PAIR_DEFER', 'd'),
        ms, 2 * fl / ms / 1e9), flush=True)


def check_pair_res32_alone():
    """RESIDUAL32 alone with parts of its epilogue disabled (DSEN2_PAIR_DEBUG bits 4/8/16), large and L2-sized batch."""
    import torch
    from dsen2_b200 import _capi
    lib, ptr = _capi.lib(), _capi.ptr
    st = _capi.stream_ptr()
    F, P = 128, 128
    for n in [int(v) for v in os.environ.get('DSEN2_DIAG_N', '84,4').split(',')]:
        w = (torch.rand((9, F, F), device='cuda') - 0.5).half()
        b = torch.zeros(F, device='cuda')
        t = torch.randn((n, P, P, F), device='cuda').half()
        hi = torch.zeros_like(t)
        x32 = torch.zeros((n, P, P // 8, F // 4, 8, 4), device='cuda')
        ms = _time(torch, lambda: _capi.check(lib.dsen2_conv_res32(ptr(t), ptr(w), ptr(b), n, P, P, 0.1, ptr(x32), ptr(hi),
                                                                   None, st), 'res32'), iters=20)
        print('DEBUG=%-2s n=%d RESIDUAL32: %.4f ms  %.2f us/patch' % (os.environ.get('DSEN2_PAIR_DEBUG', '0'), n, ms, ms * 1e3 / n), flush=True)
        xq = torch.zeros((n, P, P // 8, F // 16, 8, 16), dtype=torch.uint8, device='cuda')
        lo = torch.zeros_like(t)
        ms = _time(torch, lambda: _capi.check(lib.dsen2_conv_resq(ptr(t), ptr(w), ptr(b), n, P, P, 0.1, ptr(hi), ptr(xq),
                                                                  None, st), 'resq'), iters=20)
        print('DEBUG=%-2s n=%d RESIDUALQ : %.4f ms  %.2f us/patch' % (os.environ.get('DSEN2_PAIR_DEBUG', '0'), n, ms, ms * 1e3 / n), flush=True)
        ms = _time(torch, lambda: _capi.check(lib.dsen2_conv_resq(ptr(t), ptr(w), ptr(b), n, P, P, 0.1, ptr(hi), ptr(xq),
                                                                  ptr(lo), st), 'resq'), iters=20)
        print('DEBUG=%-2s n=%d RESIDUALQ last: %.4f ms  %.2f us/patch' % (os.environ.get('DSEN2_PAIR_DEBUG', '0'), n, ms, ms * 1e3 / n), flush=True)
        ms = _time(torch, lambda: _capi.check(lib.dsen2_conv3x3(ptr(hi), ptr(w), ptr(b), n, P, P, F, F, 9, 0, None, None,
                                                                0.0, ptr(t), None, None, None, 0, st), 'relu'), iters=20)
        print('DEBUG=%-2s n=%d RELU      : %.4f ms  %.2f us/patch' % (os.environ.get('DSEN2_PAIR_DEBUG', '0'), n, ms, ms * 1e3 / n), flush=True)


def check_pair_res32_debug_sweep():
    for dbg in os.environ.get('DSEN2_DIAG_DBG', '0,4,8,12,16,28,1,3').split(','):
        env = dict(os.environ, DSEN2_PAIR_DEBUG=dbg)
        subprocess.run([sys.executable, os.path.abspath(__file__), '--one', 'pair_res32_alone'], env=env, timeout=120)


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


def check_pair_knob_sweep():
    for pa, px, de in (('2', '1', '0'), ('2', '1', '1'), ('3', '2', '1'), ('3', '1', '0'), ('2', '2', '0'), ('1', '1', '0'),
                       ('0', '0', '0'), ('4', '1', '0'), ('2', '3', '1')):
        env = dict(os.environ, DSEN2_PAIR_PF_A=pa, DSEN2_PAIR_PF_X=px, DSEN2_PAIR_DEFER=de)
        subprocess.run([sys.executable, os.path.abspath(__file__), '--one', 'pair_res32_only'], env=env, timeout=120)


def check_pair_debug_sweep():
    for dbg in ('0', '1', '2', '3'):
        env = dict(os.environ, DSEN2_PAIR_DEBUG=dbg)
        subprocess.run([sys.executable, os.path.abspath(__file__), '--one', 'pair_relu_only'], env=env, timeout=120)


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
