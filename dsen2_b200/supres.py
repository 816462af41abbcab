"""Host-side mirror of the reference ``testing/supres.py``: ``DSen2_20`` / ``DSen2_60``.

Same signatures, ``SCALE`` / ``MDL_PATH`` module constants and weight-file naming (supres.py:11-12,57,60).
The whole tile stays on the GPU between the upload of the inputs and the download of the stitched
image.  DSen2 (128 features): per patch batch one input-preparation kernel (extract + bilinear + /2000
fused, straight from the images), 14 tcgen05 convolutions, the last one writing the stitched x2000 canvas.
VDSen2 (256 features): extract(/2000) -> bilinear(/2000) -> tcgen05 network -> stitch(x2000).
"""
import os

import numpy as np

from . import _capi, sharding
from .DSen2Net import s2model
from .patches import bilinear_up_device, extract_patches_device, patch_counts, recompose_device

SCALE = 2000
MDL_PATH = '../models/'

_GEOM = {False: dict(patch=128, border=8, ratio=2), True: dict(patch=192, border=12, ratio=6)}
_model_cache = {}


def weight_file(deep=False, run_60=False):
    """File-name convention of supres.py:57,60."""
    if deep:
        return MDL_PATH + ('s2_034_lr_1e-04.hdf5' if run_60 else 's2_033_lr_1e-04.hdf5')
    return MDL_PATH + ('s2_030_lr_1e-05.hdf5' if run_60 else 's2_032_lr_1e-04.hdf5')


def _load_model(input_shape, deep, run_60):
    path = weight_file(deep, run_60)
    st = os.stat(path)                      # FileNotFoundError (OSError) when the weights are missing, as h5py
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
    if key not in _model_cache:
        model = s2model(input_shape, num_layers=32, feature_size=256) if deep else \
            s2model(input_shape, num_layers=6, feature_size=128)
        print('Symbolic Model Created.')
        model.load_weights(path)
        _model_cache[key] = model
    print("Predicting using file: {}".format(path))
    return _model_cache[key]


def default_device_batch(W, P, B):
    """Patches per launch: whole patch rows of the tile when a row is a reasonable batch (a full Sentinel-2 tile has 99
    patches of 128 per row), so that the chunks of ``HostPipeline`` (whole rows) split into equal batches; about 4.9 M
    patch pixels otherwise.  Larger batches only amortise launch overhead and the last wave of each launch: measured on
    one box 559-561 ms per tile with 99 patches per launch, 553-555 ms with 198 or 297."""
    nx = -(-W // (P - 2 * B))
    target = max(1, (300 * 128 * 128) // (P * P))
    if nx > 2 * target:
        return target
    return nx * max(1, target // nx)


def super_resolve_device(model, d10, d20, d60=None, first_patch=0, num_patches=None, out=None, device_batch=None,
                         timers=None):
    """Device-level pipeline.  d10/d20(/d60): CUDA float32 HWC tensors.  Processes patches
    [first_patch, first_patch+num_patches) of the FILLED patch list (surplus zero patches of
    patches.py:32-39 are never read by recompose_images, so they are skipped) and writes the pixels those
    patches own into ``out`` (H, W, Cout) float32 (allocated zero-filled if None)."""
    torch = _capi.require_cuda()
    run_60 = d60 is not None
    g = _GEOM[run_60]
    r, P, B = g['ratio'], g['patch'], g['border']
    H, W = int(d10.shape[0]), int(d10.shape[1])
    if H % r or W % r:
        raise ValueError("10 m image size %dx%d must be a multiple of %d" % (H, W, r))
    plr, blr = P // r, B // r
    _, filled = patch_counts(H // r, W // r, plr, blr)
    if num_patches is None:
        num_patches = filled - first_patch
    if first_patch < 0 or first_patch + num_patches > filled:
        raise ValueError("patch range [%d, %d) outside the %d patches of this tile"
                         % (first_patch, first_patch + num_patches, filled))
    if out is None:
        out = torch.zeros((H, W, model.out_channels), dtype=torch.float32, device=d10.device)
    if device_batch is None:
        device_batch = default_device_batch(W, P, B)
    single = filled == 1                     # recompose_images returns the lone patch uncropped (patches.py:375-376)
    for p0 in range(first_patch, first_patch + num_patches, device_batch):
        nb = min(device_batch, first_patch + num_patches - p0)
        if model.fast_path and not single:       # fused: images -> x_in -> network -> stitched canvas
            model.forward_images(d10, d20, d60, P, B, p0, nb, out, float(SCALE), timers=timers)
            continue
        if run_60:
            xs = [extract_patches_device(d10, 6, plr, blr, p0, nb, divisor=SCALE),
                  bilinear_up_device(extract_patches_device(d20, 3, plr, blr, p0, nb), 2, post_divisor=SCALE),
                  bilinear_up_device(extract_patches_device(d60, 1, plr, blr, p0, nb), 6, post_divisor=SCALE)]
        else:
            xs = [extract_patches_device(d10, 2, plr, blr, p0, nb, divisor=SCALE),
                  bilinear_up_device(extract_patches_device(d20, 1, plr, blr, p0, nb), 2, post_divisor=SCALE)]
        pred = model.forward_device(xs, timers=timers)
        if single:
            return (pred[0] * float(SCALE)).permute(1, 2, 0).contiguous()
        recompose_device(pred, B, H, W, first_patch=p0, mul=float(SCALE), out=out)
    return out


class HostPipeline:
    """Host-buffer front end: tile inputs / output live in (pinned) HOST memory; the patch range is cut into
    chunks of whole patch rows and chunk k+1's input rows upload while chunk k computes and chunk k-1's owned
    output rows download (three CUDA streams, one device-resident tile).  Reusable across calls of one shape."""

    def __init__(self, model, H, W, run_60=False, device=None, chunk_patch_rows=None, device_batch=None):
        torch = _capi.require_cuda()
        self.torch, self.model, self.run_60 = torch, model, run_60
        self.H, self.W = int(H), int(W)
        g = _GEOM[run_60]
        self.r, self.P, self.B = g['ratio'], g['patch'], g['border']
        if self.H % self.r or self.W % self.r:
            raise ValueError("10 m image size %dx%d must be a multiple of %d" % (H, W, self.r))
        self.dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        self.ny, self.nx, self.S = sharding.tile_grid(self.H, self.W, self.P, self.B)
        self.chunk_rows, self.device_batch = chunk_patch_rows, device_batch
        mk = lambda h, w, c: torch.empty((h, w, c), dtype=torch.float32, device=self.dev)
        self.d10, self.d20 = mk(self.H, self.W, 4), mk(self.H // 2, self.W // 2, 6)
        self.d60 = mk(self.H // 6, self.W // 6, 2) if run_60 else None
        self.canvas = torch.zeros((self.H, self.W, model.out_channels), dtype=torch.float32, device=self.dev)
        self.up, self.down = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.h2d_bytes = self.d2h_bytes = 0

    def _upload(self, h, d, div, r0, r1, done):
        """rows [r0, r1) of the 10 m grid -> rows of the (H/div) grid not uploaded yet."""
        a, b = r0 // div, -(-r1 // div)
        if done is not None:
            a = max(a, done)
        if b > a:
            d[a:b].copy_(h[a:b], non_blocking=True)
            self.h2d_bytes += (b - a) * d.shape[1] * d.shape[2] * 4
        return b if done is None else max(done, b)

    def run(self, h10, h20, h60=None, hout=None, first_patch=0, num_patches=None, timers=None):
        torch = self.torch
        filled = self.ny * self.nx
        if num_patches is None:
            num_patches = filled - first_patch
        if hout is None:
            hout = torch.empty((self.H, self.W, self.model.out_channels), dtype=torch.float32).pin_memory()
        self.h2d_bytes = self.d2h_bytes = 0
        main = torch.cuda.current_stream(self.dev)
        self.up.wait_stream(main)
        self.down.wait_stream(main)
        done10 = done20 = done60 = None           # rows already resident on the device (per resolution)
        chunk_rows = self.chunk_rows
        if chunk_rows is None:
            chunk_rows = sharding.auto_chunk_rows(num_patches, self.nx)
        for p0, cnt, (r0, r1), rects in sharding.plan_chunks(first_patch, num_patches, self.H, self.W, self.P, self.B,
                                                             int(chunk_rows)):
            with torch.cuda.stream(self.up):
                done10 = self._upload(h10, self.d10, 1, r0, r1, done10)
                done20 = self._upload(h20, self.d20, 2, r0, r1, done20)
                if self.run_60:
                    done60 = self._upload(h60, self.d60, 6, r0, r1, done60)
                ev_up = torch.cuda.Event()
                ev_up.record(self.up)
            main.wait_event(ev_up)
            super_resolve_device(self.model, self.d10, self.d20, self.d60, first_patch=p0, num_patches=cnt,
                                 out=self.canvas, device_batch=self.device_batch, timers=timers)
            ev_c = torch.cuda.Event()
            ev_c.record(main)
            self.down.wait_event(ev_c)
            with torch.cuda.stream(self.down):
                for (y0, y1, x0, x1) in rects:
                    if x0 == 0 and x1 == self.W:
                        hout[y0:y1].copy_(self.canvas[y0:y1], non_blocking=True)
                    else:
                        hout[y0:y1, x0:x1].copy_(self.canvas[y0:y1, x0:x1], non_blocking=True)
                    self.d2h_bytes += (y1 - y0) * (x1 - x0) * self.canvas.shape[2] * 4
        main.wait_stream(self.down)
        return hout


def _run(model, arrays):
    torch = _capi.require_cuda()
    host = [torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.float32)) for a in arrays]
    H, W = int(host[0].shape[0]), int(host[0].shape[1])
    run_60 = len(host) == 3
    g = _GEOM[run_60]
    ny, nx, _ = sharding.tile_grid(H, W, g['patch'], g['border'])
    if ny * nx == 1 or not model.fast_path:      # single patch (returned uncropped) / VDSen2: plain device path
        out = super_resolve_device(model, *[t.cuda() for t in host])
        return out.cpu().numpy()
    pipe = HostPipeline(model, H, W, run_60=run_60)
    hout = pipe.run(*host)
    torch.cuda.current_stream().synchronize()
    return hout.numpy()


def DSen2_20(d10, d20, deep=False, model=None):
    """supres.py:15-30.  d10 (H,W,4), d20 (H/2,W/2,6) -> (H,W,6) float32.  ``model`` (extension) supplies a
    preloaded ``S2Model`` instead of the shipped hdf5 weights."""
    input_shape = ((4, None, None), (6, None, None))
    if model is None:
        model = _load_model(input_shape, deep, run_60=False)
    return _run(model, [d10, d20])


def DSen2_60(d10, d20, d60, deep=False, model=None):
    """supres.py:33-50.  + d60 (H/6,W/6,2) -> (H,W,2) float32."""
    input_shape = ((4, None, None), (6, None, None), (2, None, None))
    if model is None:
        model = _load_model(input_shape, deep, run_60=True)
    return _run(model, [d10, d20, d60])
