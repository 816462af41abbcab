"""Host-side mirror of the reference ``utils/imresize.py`` (MATLAB-compatible bicubic).

``imresize(I, scalar_scale=None, output_shape=None)`` as imresize.py:80-112.  The tap tables are
built on the host in float64 (a few KB; same formulas as ``contributions`` :28-48) and the
separable filter runs in ``bicubic_tiled_kernel`` (float64, left-to-right sums => bit-identical).
"""
from math import ceil

import numpy as np

from . import _capi


def _cubic(x):
    ax = np.absolute(np.asarray(x, np.float64))
    ax2 = np.multiply(ax, ax)
    ax3 = np.multiply(ax2, ax)
    return np.multiply(1.5 * ax3 - 2.5 * ax2 + 1, ax <= 1) + \
        np.multiply(-0.5 * ax3 + 2.5 * ax2 - 4 * ax + 2, (1 < ax) & (ax <= 2))


def tap_tables(in_length, out_length, scale, kernel_width=4.0):
    """(weights float64 (out, T), indices int32 (out, T)) -- imresize.py:28-48."""
    if scale < 1:
        h = lambda t: scale * _cubic(scale * t)
        kw = 1.0 * kernel_width / scale
    else:
        h, kw = _cubic, kernel_width
    x = np.arange(1, out_length + 1).astype(np.float64)
    u = x / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - kw / 2)
    P = int(ceil(kw)) + 2
    indices = (np.expand_dims(left, axis=1) + np.arange(P) - 1).astype(np.int32)
    weights = h(np.expand_dims(u, axis=1) - indices - 1)
    weights = np.divide(weights, np.expand_dims(np.sum(weights, axis=1), axis=1))
    aux = np.concatenate((np.arange(in_length), np.arange(in_length - 1, -1, step=-1))).astype(np.int32)
    indices = aux[np.mod(indices, aux.size)]
    keep = np.nonzero(np.any(weights, axis=0))[0]
    return np.ascontiguousarray(weights[:, keep]), np.ascontiguousarray(indices[:, keep].astype(np.int32))


_TABLES = {}


def _device_tables(dev, in_length, out_length, scale):
    """Tap tables of one dimension on ``dev`` (cached: a tile is resized with the same tables every time)."""
    key = (str(dev), in_length, out_length, scale)
    hit = _TABLES.get(key)
    if hit is None:
        torch = _capi.require_cuda()
        if len(_TABLES) > 32:
            _TABLES.clear()
        w, i = tap_tables(in_length, out_length, scale)
        hit = _TABLES[key] = (torch.from_numpy(w).to(dev), torch.from_numpy(i).to(dev), int(w.shape[1]))
    return hit


def imresize_device(img, scale, output_size):
    """img: CUDA float32/float64 (h, w, C) -> CUDA float64 (out_h, out_w, C)."""
    torch = _capi.require_cuda()
    assert img.is_cuda and img.is_contiguous() and img.dim() == 3 and img.dtype in (torch.float32, torch.float64)
    h, w, C = img.shape
    dev = img.device
    wy, iy, ty_ = _device_tables(dev, h, int(output_size[0]), float(scale[0]))
    wx, ix, tx_ = _device_tables(dev, w, int(output_size[1]), float(scale[1]))
    order = np.argsort(np.array(scale))           # imresize.py:95 (stable for equal scales: dim 0 first)
    t = [wy, iy, wx, ix]
    out = torch.empty((output_size[0], output_size[1], C), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = _capi.lib().dsen2_bicubic_imresize(_capi.ptr(img), int(img.dtype == torch.float64), h, w, C,
                                                _capi.ptr(t[0]), _capi.ptr(t[1]), ty_, output_size[0],
                                                _capi.ptr(t[2]), _capi.ptr(t[3]), tx_, output_size[1],
                                                int(order[0]), _capi.ptr(out), _capi.stream_ptr())
    _capi.check(rc, "dsen2_bicubic_imresize")
    return out


def imresize(I, scalar_scale=None, output_shape=None):
    I = np.asarray(I)
    if scalar_scale is not None:
        scalar_scale = float(scalar_scale)
        scale = [scalar_scale, scalar_scale]
        output_size = [int(ceil(scale[k] * I.shape[k])) for k in range(2)]
    elif output_shape is not None:
        scale = [1.0 * output_shape[k] / I.shape[k] for k in range(2)]
        output_size = list(output_shape)
    else:
        print('Error: scalar_scale OR output_shape should be defined!')   # imresize.py:91-93
        return
    if I.dtype == np.uint8:
        raise NotImplementedError("uint8 rounding path (imresize.py:70-72) is not on the DSen2 hot path")
    torch = _capi.require_cuda()
    B = I if I.ndim == 3 else I[:, :, None]
    dt = np.float64 if B.dtype == np.float64 else np.float32
    out = imresize_device(torch.from_numpy(np.ascontiguousarray(B, dtype=dt)).cuda(), scale, output_size).cpu().numpy()
    return out if I.ndim == 3 else out[:, :, 0]
