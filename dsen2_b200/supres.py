"""Host-side mirror of the reference ``testing/supres.py``: ``DSen2_20`` / ``DSen2_60``.

Same signatures, ``SCALE`` / ``MDL_PATH`` module constants and weight-file naming (supres.py:11-12,57,60).
The whole tile stays on the GPU between the upload of the inputs and the download of the stitched
image.  Per patch batch: one input-preparation kernel (extract + bilinear + /2000 fused, straight from the images),
then the 14 (DSen2, 6 x 128) or 66 (VDSen2, 32 x 256) tcgen05 convolutions, the last one writing the stitched
x2000 canvas.
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
    """Device-level pipeline.  d10/d20(/d60): CUDA HWC tensors, all float32 or all uint16.  Processes patches
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
    # recompose_images' "lone patch is returned uncropped" branch (patches.py:375-376) keys on the ALLOCATED patch count
    # (k_i+1)(k_j+1), which is 1 only for images smaller than one stride -- inputs get_test_patches itself cannot tile
    # (and dsen2_prep16_from_images rejects); every image that reaches this point is stitched and cropped.
    for p0 in range(first_patch, first_patch + num_patches, device_batch):
        nb = min(device_batch, first_patch + num_patches - p0)
        if model.xin16:                          # fused: images -> x_in16 -> network -> stitched canvas
            model.forward_images(d10, d20, d60, P, B, p0, nb, out, float(SCALE), timers=timers)
            continue
        # a network without resblocks: separate extract / bilinear / network / stitch kernels
        if run_60:
            xs = [extract_patches_device(d10, 6, plr, blr, p0, nb, divisor=SCALE),
                  bilinear_up_device(extract_patches_device(d20, 3, plr, blr, p0, nb), 2, post_divisor=SCALE),
                  bilinear_up_device(extract_patches_device(d60, 1, plr, blr, p0, nb), 6, post_divisor=SCALE)]
        else:
            xs = [extract_patches_device(d10, 2, plr, blr, p0, nb, divisor=SCALE),
                  bilinear_up_device(extract_patches_device(d20, 1, plr, blr, p0, nb), 2, post_divisor=SCALE)]
        pred = model.forward_device(xs, timers=timers)
        recompose_device(pred, B, H, W, first_patch=p0, mul=float(SCALE), out=out)
    return out


def _np_dtype_of(torch, dt):
    return {torch.float32: np.float32, torch.uint16: np.uint16}[dt]


class _Slot:
    """One slot of the pinned staging ring: input rows of a chunk per resolution + the output rectangles it owns."""

    def __init__(self, torch, shapes_in, dtype, out_elems):
        self.inp = [torch.empty(s, dtype=dtype).pin_memory() for s in shapes_in]
        self.out = torch.empty((out_elems,), dtype=torch.float32).pin_memory()
        self.ev_h2d, self.ev_d2h = torch.cuda.Event(), torch.cuda.Event()
        self.used_in = self.used_out = False


class HostPipeline:
    """Host-buffer front end: tile inputs / output live in HOST memory; the patch range is cut into chunks of whole
    patch rows and chunk k+1's input rows upload while chunk k computes and chunk k-1's owned output rows download
    (three CUDA streams, one device-resident tile).  Reusable across calls of one shape.

    Host buffers may be pinned torch tensors (copied directly, fully asynchronous) or ordinary pageable numpy arrays
    -- what a caller of ``DSen2_20`` holds.  Pageable arrays go through a ring of pinned staging slots filled / drained
    by worker threads (numpy releases the GIL for the copies), so neither the page-locked upload nor the first-touch
    page faults of a freshly allocated output serialise with the GPU.  ``dtype``: element type of the device-resident
    images, float32 or uint16 (Sentinel-2 digital numbers as GDAL delivers them, s2_tiles_supres.py:311-315: half the
    bytes over PCIe; ``dsen2_prep16_from_images`` converts exactly, results are bit-identical)."""

    RING = 3

    def __init__(self, model, H, W, run_60=False, device=None, chunk_patch_rows=None, device_batch=None, dtype=None):
        torch = _capi.require_cuda()
        self.torch, self.model, self.run_60 = torch, model, run_60
        self.H, self.W = int(H), int(W)
        g = _GEOM[run_60]
        self.r, self.P, self.B = g['ratio'], g['patch'], g['border']
        if self.H % self.r or self.W % self.r:
            raise ValueError("10 m image size %dx%d must be a multiple of %d" % (H, W, self.r))
        self.dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        self.dtype = torch.float32 if dtype is None else dtype
        if self.dtype not in (torch.float32, torch.uint16):
            raise ValueError("device images are float32 or uint16")
        self.ny, self.nx, self.S = sharding.tile_grid(self.H, self.W, self.P, self.B)
        self.chunk_rows, self.device_batch = chunk_patch_rows, device_batch
        mk = lambda h, w, c: torch.empty((h, w, c), dtype=self.dtype, device=self.dev)
        self.d10, self.d20 = mk(self.H, self.W, 4), mk(self.H // 2, self.W // 2, 6)
        self.d60 = mk(self.H // 6, self.W // 6, 2) if run_60 else None
        self.canvas = torch.zeros((self.H, self.W, model.out_channels), dtype=torch.float32, device=self.dev)
        self.up, self.down = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.h2d_bytes = self.d2h_bytes = 0
        self._slots, self._pool = None, None

    # ---- pinned host buffers: direct asynchronous copies ------------------------------------------------------------
    def _upload(self, h, d, div, r0, r1, done):
        """rows [r0, r1) of the 10 m grid -> rows of the (H/div) grid not uploaded yet."""
        a, b = r0 // div, -(-r1 // div)
        if done is not None:
            a = max(a, done)
        if b > a:
            d[a:b].copy_(h[a:b], non_blocking=True)
            self.h2d_bytes += (b - a) * d.shape[1] * d.shape[2] * d.element_size()
        return b if done is None else max(done, b)

    def run(self, h10, h20, h60=None, hout=None, first_patch=0, num_patches=None, timers=None):
        """Pinned torch tensors in (dtype = the pipeline's), pinned float32 (H, W, Cout) tensor out.  Rows of ``hout`` that
        the patch range owns only partly (the seam rows of a sharded range) are downloaded whole."""
        torch = self.torch
        filled = self.ny * self.nx
        if num_patches is None:
            num_patches = filled - first_patch
        if hout is None:
            hout = torch.empty((self.H, self.W, self.model.out_channels), dtype=torch.float32).pin_memory()
        for h in (h10, h20) + ((h60,) if self.run_60 else ()):
            if h.dtype != self.dtype:
                raise ValueError("host image dtype %s does not match the pipeline's %s" % (h.dtype, self.dtype))
        self.h2d_bytes = self.d2h_bytes = 0
        main = torch.cuda.current_stream(self.dev)
        self.up.wait_stream(main)
        self.down.wait_stream(main)
        done10 = done20 = done60 = None           # rows already resident on the device (per resolution)
        for p0, cnt, (r0, r1), rects in self._plan(first_patch, num_patches):
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
                    # Whole rows, also where the range starts / ends inside a patch row and owns only part of them: a
                    # strided (partial-width) device -> host copy goes through temporaries and synchronises.  The pixels
                    # of those rows that another rank owns come along as they are on this device (never computed here);
                    # `sharding.assemble` / `owned_rects` say which part of a row is this rank's.
                    hout[y0:y1].copy_(self.canvas[y0:y1], non_blocking=True)
                    self.d2h_bytes += (y1 - y0) * self.W * self.canvas.shape[2] * 4
        main.wait_stream(self.down)
        return hout

    def _plan(self, first_patch, num_patches):
        if self.chunk_rows is None:
            batch = self.device_batch or default_device_batch(self.W, self.P, self.B)
            middle, edge = sharding.auto_chunk_rows(num_patches, self.nx, batch)
        else:
            middle = edge = int(self.chunk_rows)
        return sharding.plan_chunks(first_patch, num_patches, self.H, self.W, self.P, self.B, middle, edge)

    # ---- pageable numpy arrays: pinned staging ring + worker threads -------------------------------------------------
    def _ring(self, plan):
        """Slots sized for the largest chunk of ``plan`` (allocated once per pipeline, grown on demand)."""
        torch = self.torch
        divs = (1, 2, 6) if self.run_60 else (1, 2)
        devs = (self.d10, self.d20, self.d60)
        rows = [max(-(-r1 // dv) - r0 // dv for _, _, (r0, r1), _ in plan) for dv in divs]
        shapes = [(rows[i], devs[i].shape[1], devs[i].shape[2]) for i in range(len(divs))]
        out_elems = max(sum((y1 - y0) * (x1 - x0) for (y0, y1, x0, x1) in rects) for _, _, _, rects in plan) \
            * self.canvas.shape[2]
        if self._slots is None or any(a.shape[0] < s[0] for a, s in zip(self._slots[0].inp, shapes)) \
                or self._slots[0].out.numel() < out_elems:
            self._slots = [_Slot(torch, shapes, self.dtype, out_elems) for _ in range(self.RING)]
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=4, thread_name_prefix='dsen2-stage',
                                            initializer=torch.cuda.set_device, initargs=(self.dev,))
        return self._slots

    def run_numpy(self, a10, a20, a60=None, out=None, first_patch=0, num_patches=None, timers=None):
        """Pageable numpy arrays in -> numpy (H, W, Cout) float32 out (allocated if None).  Arrays whose dtype is not
        the pipeline's are cast while they are staged."""
        torch = self.torch
        filled = self.ny * self.nx
        if num_patches is None:
            num_patches = filled - first_patch
        C = self.model.out_channels
        if out is None:
            out = np.empty((self.H, self.W, C), np.float32)
        srcs = [a10, a20] + ([a60] if self.run_60 else [])
        divs = (1, 2, 6) if self.run_60 else (1, 2)
        devs = [self.d10, self.d20] + ([self.d60] if self.run_60 else [])
        plan = self._plan(first_patch, num_patches)
        slots = self._ring(plan)
        # rows of every resolution a chunk has to bring (those beyond what earlier chunks brought)
        need, done = [], [None] * len(divs)
        for _, _, (r0, r1), _ in plan:
            rows = []
            for i, dv in enumerate(divs):
                a, b = r0 // dv, -(-r1 // dv)
                if done[i] is not None:
                    a = max(a, done[i])
                rows.append((a, max(a, b)))
                done[i] = max(a, b) if done[i] is None else max(done[i], b)
            need.append(rows)
        self.h2d_bytes = self.d2h_bytes = 0

        def stage_in(k):
            slot = slots[k % self.RING]
            if slot.used_in:
                slot.ev_h2d.synchronize()            # the previous upload from this slot has left it
            for i, (a, b) in enumerate(need[k]):
                if b > a:
                    slot.inp[i][:b - a].numpy()[...] = srcs[i][a:b]

        def stage_out(rects, src, offs, part, parts):
            for (y0, y1, x0, x1), o in zip(rects, offs):
                h = y1 - y0
                ya, yb = y0 + h * part // parts, y0 + h * (part + 1) // parts
                if yb > ya:
                    view = src[o:o + h * (x1 - x0) * C].reshape(h, x1 - x0, C)
                    out[ya:yb, x0:x1] = view[ya - y0:yb - y0]

        main = torch.cuda.current_stream(self.dev)
        self.up.wait_stream(main)
        self.down.wait_stream(main)
        fin = {k: self._pool.submit(stage_in, k) for k in range(min(self.RING, len(plan)))}
        fout = {}
        for k, (p0, cnt, _rows, rects) in enumerate(plan):
            slot = slots[k % self.RING]
            fin.pop(k).result()
            with torch.cuda.stream(self.up):
                for i, (a, b) in enumerate(need[k]):
                    if b > a:
                        devs[i][a:b].copy_(slot.inp[i][:b - a], non_blocking=True)
                        self.h2d_bytes += (b - a) * devs[i].shape[1] * devs[i].shape[2] * devs[i].element_size()
                slot.ev_h2d.record(self.up)
                slot.used_in = True
            main.wait_event(slot.ev_h2d)
            if k + self.RING < len(plan):
                fin[k + self.RING] = self._pool.submit(stage_in, k + self.RING)
            super_resolve_device(self.model, self.d10, self.d20, self.d60, first_patch=p0, num_patches=cnt,
                                 out=self.canvas, device_batch=self.device_batch, timers=timers)
            ev_c = torch.cuda.Event()
            ev_c.record(main)
            self.down.wait_event(ev_c)
            for f in fout.pop(k - self.RING, ()):    # the slot's previous contents have been copied out
                f.result()
            offs, o = [], 0
            with torch.cuda.stream(self.down):
                for (y0, y1, x0, x1) in rects:
                    nel = (y1 - y0) * (x1 - x0) * C
                    dst = slot.out[o:o + nel].view(y1 - y0, x1 - x0, C)
                    dst.copy_(self.canvas[y0:y1] if (x0 == 0 and x1 == self.W) else self.canvas[y0:y1, x0:x1],
                              non_blocking=True)
                    offs.append(o)
                    o += nel
                    self.d2h_bytes += nel * 4
                slot.ev_d2h.record(self.down)
            src = slot.out.numpy()

            def drain(part, parts, rects=rects, src=src, offs=offs, ev=slot.ev_d2h):
                ev.synchronize()
                stage_out(rects, src, offs, part, parts)
            fout[k] = [self._pool.submit(drain, part, 2) for part in range(2)]
        for fs in fout.values():
            for f in fs:
                f.result()
        main.wait_stream(self.down)
        return out


_pipe_cache = {}


def _pipeline_for(model, H, W, run_60, dtype):
    """HostPipeline objects (device-resident tile + pinned staging ring) are kept for the last shapes used, so a caller
    that super-resolves scene after scene pays the allocations once."""
    torch = _capi.require_cuda()
    key = (id(model), H, W, run_60, dtype, torch.cuda.current_device())
    pipe = _pipe_cache.get(key)
    if pipe is None or pipe.model is not model:
        if len(_pipe_cache) >= 2:
            _pipe_cache.clear()
        pipe = _pipe_cache[key] = HostPipeline(model, H, W, run_60=run_60, dtype=dtype)
    return pipe


def _run(model, arrays, out=None):
    torch = _capi.require_cuda()
    arrays = [np.asarray(a) for a in arrays]
    H, W = int(arrays[0].shape[0]), int(arrays[0].shape[1])
    for a, div, c in zip(arrays, (1, 2, 6), model.in_channels):
        if a.ndim != 3 or tuple(a.shape) != (H // div, W // div, c):
            raise ValueError("expected an image of shape %s (HWC, 10 m size %dx%d / %d), got %s"
                             % ((H // div, W // div, c), H, W, div, tuple(a.shape)))
    # uint16 digital numbers (GDAL, s2_tiles_supres.py:311-315) stay uint16 up to the input-preparation kernel; everything
    # else is staged as float32 (the reference divides by SCALE in floating point whatever the input type, supres.py:23-24)
    dtype = torch.uint16 if (model.xin16 and all(a.dtype == np.uint16 for a in arrays)) else torch.float32
    pipe = _pipeline_for(model, H, W, len(arrays) == 3, dtype)
    res = pipe.run_numpy(*arrays, out=out)
    torch.cuda.current_stream().synchronize()
    return res


def DSen2_20(d10, d20, deep=False, model=None, out=None):
    """supres.py:15-30.  d10 (H,W,4), d20 (H/2,W/2,6) -> (H,W,6) float32.  Extensions: ``model`` supplies a preloaded
    ``S2Model`` instead of the shipped hdf5 weights, ``out`` a preallocated (H,W,6) float32 array to fill."""
    input_shape = ((4, None, None), (6, None, None))
    if model is None:
        model = _load_model(input_shape, deep, run_60=False)
    return _run(model, [d10, d20], out)


def DSen2_60(d10, d20, d60, deep=False, model=None, out=None):
    """supres.py:33-50.  + d60 (H/6,W/6,2) -> (H,W,2) float32."""
    input_shape = ((4, None, None), (6, None, None), (2, None, None))
    if model is None:
        model = _load_model(input_shape, deep, run_60=True)
    return _run(model, [d10, d20, d60], out)
