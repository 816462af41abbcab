"""CPU restatement of the reference tiling / stitching / bilinear upsample.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows
``/root/reference/utils/patches.py``:

* ``get_test_patches``    -> patches.py:19-80
* ``get_test_patches60``  -> patches.py:83-156
* ``interp_patches``      -> patches.py:11-16 (skimage ``resize(mode='reflect')``, order 1)
* ``recompose_images``    -> patches.py:374-405

Written from the index formulas, not from the reference's
loops: one generic multi-resolution extractor serves both entry points.
"""
import numpy as np


def sym_index(j, n):
    """numpy ``pad(mode='symmetric')`` source index for padded coordinate j (patches.py:27-28)."""
    j = np.asarray(j)
    j = np.where(j < 0, -j - 1, j)
    return np.where(j >= n, 2 * n - 1 - j, j)


def axis_starts(n_lr, patch_lr, border_lr):
    """Crop starts on the tiling grid, in padded low-res coordinates (patches.py:45-53)."""
    stride = patch_lr - 2 * border_lr
    k = n_lr // stride
    starts = [t * stride for t in range(k)]
    if n_lr % stride != 0:
        starts.append(n_lr + 2 * border_lr - patch_lr)
    return starts, k + 1  # (filled starts, allocated count) -- patches.py:32-36


def _extract(dsets, ratios, patch_lr, border_lr):
    """dsets[i] is (H*r_i, W*r_i, C_i) where (H, W) is the tiling-grid size and r_i = ratios[i]."""
    H, W = dsets[-1].shape[0] // ratios[-1], dsets[-1].shape[1] // ratios[-1]
    si, ai = axis_starts(H, patch_lr, border_lr)
    sj, aj = axis_starts(W, patch_lr, border_lr)
    outs = []
    for d, r in zip(dsets, ratios):
        d = np.asarray(d)
        p, b = patch_lr * r, border_lr * r
        out = np.zeros((ai * aj, d.shape[2], p, p), np.float32)
        n = 0
        for ii in si:
            ys = sym_index(np.arange(ii * r, ii * r + p) - b, d.shape[0])
            for jj in sj:
                xs = sym_index(np.arange(jj * r, jj * r + p) - b, d.shape[1])
                out[n] = d[np.ix_(ys, xs)].transpose(2, 0, 1)
                n += 1
        outs.append(out)
    return outs


def interp_patches(image_lr, image_hr_shape):
    """Bilinear upsample of every (patch, band) with mirror boundary (patches.py:11-16).

    skimage >= 0.19 ``resize(x, shape, mode='reflect')`` (order 1, upscaling => no
    anti-aliasing) == ``scipy.ndimage.zoom(x, s, order=1, mode='mirror', grid_mode=True)``:
    sample position u = (o + 0.5)/s - 0.5, mirror index -1 -> 1, n -> n-2.  The
    reference divides by 30000 before and multiplies after (float32 array ops).
    """
    n, c, p, _ = image_lr.shape
    P = image_hr_shape[2]
    s = P // p
    o = np.arange(P)
    u = (o + 0.5) / s - 0.5
    i0 = np.floor(u).astype(np.int64)
    f = u - i0

    def mir(i):
        i = np.where(i < 0, -i, i)
        return np.where(i > p - 1, 2 * (p - 1) - i, i)

    a0, a1 = mir(i0), mir(i0 + 1)
    x = (image_lr.astype(np.float32) / np.float32(30000)).astype(np.float64)
    rows = x[:, :, a0, :] * (1 - f)[None, None, :, None] + x[:, :, a1, :] * f[None, None, :, None]
    out = rows[:, :, :, a0] * (1 - f) + rows[:, :, :, a1] * f
    return (out.astype(np.float32) * np.float32(30000)).astype(np.float32)


def get_test_patches(dset_10, dset_20, patchSize=128, border=4, interp=True):
    p10, p20 = _extract([dset_10, dset_20], [2, 1], patchSize // 2, border // 2)
    if interp:
        p20 = interp_patches(p20, p10.shape)
    return p10, p20


def get_test_patches60(dset_10, dset_20, dset_60, patchSize=128, border=8, interp=True):
    p10, p20, p60 = _extract([dset_10, dset_20, dset_60], [6, 3, 1], patchSize // 6, border // 6)
    if interp:
        p20 = interp_patches(p20, p10.shape)
        p60 = interp_patches(p60, p10.shape)
    return p10, p20, p60


def stitch_tile_of(y, size, S):
    """Index of the tile that wrote output coordinate y last (patches.py:394-403)."""
    n = -(-size // S)
    y = np.asarray(y)
    t = y // S
    if size % S != 0:
        t = np.where(y >= size - S, n - 1, t)
    return t, n


def recompose_images(a, border, size=None):
    """Stitch predicted patches (patches.py:374-405) as a gather: last writer wins."""
    if a.shape[0] == 1:
        images = a[0]
    else:
        S = a.shape[2] - 2 * border
        H, W = size[0], size[1]
        ys, xs = np.arange(H), np.arange(W)
        ty, _ny = stitch_tile_of(ys, H, S)
        tx, nx = stitch_tile_of(xs, W, S)
        oy = np.minimum(ty * S, H - S)
        ox = np.minimum(tx * S, W - S)
        p = ty[:, None] * nx + tx[None, :]
        images = a[p, :, (border + ys - oy)[:, None], (border + xs - ox)[None, :]]  # (H, W, C)
        images = np.ascontiguousarray(images.transpose(2, 0, 1)).astype(np.float32)
    return images.transpose((1, 2, 0))


def downPixelAggr(img, SCALE=2):
    """patches.py:353-371: Gaussian blur (sigma = 1/SCALE, per band) then SCALE x SCALE pixel aggregation (block mean).

    Restated with ``scipy.ndimage.gaussian_filter`` (the reference's own call) and a reshape-mean in place of
    ``skimage.measure.block_reduce`` (scikit-image is absent, so the reference function itself cannot run here:
    parity unpinned by the reference, pinned against scipy)."""
    from scipy.ndimage import gaussian_filter
    img = np.asarray(img)
    if img.ndim == 2:
        img = img[:, :, None]
    blur = np.zeros(img.shape)
    for i in range(img.shape[2]):
        blur[:, :, i] = gaussian_filter(img[:, :, i], 1 / SCALE)
    h, w = img.shape[0] // SCALE, img.shape[1] // SCALE
    lr = blur[:h * SCALE, :w * SCALE].reshape(h, SCALE, w, SCALE, img.shape[2]).transpose(0, 2, 4, 1, 3)
    return np.squeeze(lr.reshape(h, w, img.shape[2], SCALE * SCALE).mean(axis=-1))
