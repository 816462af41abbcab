"""CPU restatement of the reference MATLAB-compatible bicubic ``imresize``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows
``/root/reference/utils/imresize.py``: ``cubic`` :20-26, ``contributions`` :28-48,
``imresizemex`` :50-74, ``imresize`` :80-112.  Vectorised; the per-tap products are
summed left to right in float64 exactly as ``np.sum`` does for <8 terms, so the
result is bit-identical to the reference's interpreted double loop.
"""
from math import ceil

import numpy as np


def cubic(x):
    """Keys kernel, a = -0.5 (imresize.py:20-26)."""
    x = np.asarray(x, np.float64)
    ax = np.absolute(x)
    ax2 = np.multiply(ax, ax)
    ax3 = np.multiply(ax2, ax)
    return np.multiply(1.5 * ax3 - 2.5 * ax2 + 1, ax <= 1) + \
        np.multiply(-0.5 * ax3 + 2.5 * ax2 - 4 * ax + 2, (1 < ax) & (ax <= 2))


def contributions(in_length, out_length, scale, k_width=4.0):
    """Tap weights (out, T) float64 and source indices (out, T) int32 (imresize.py:28-48)."""
    if scale < 1:
        h = lambda t: scale * cubic(scale * t)
        kernel_width = 1.0 * k_width / scale
    else:
        h = cubic
        kernel_width = k_width
    x = np.arange(1, out_length + 1).astype(np.float64)
    u = x / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - kernel_width / 2)
    P = int(ceil(kernel_width)) + 2
    ind = np.expand_dims(left, axis=1) + np.arange(P) - 1
    indices = ind.astype(np.int32)
    weights = h(np.expand_dims(u, axis=1) - indices - 1)
    weights = np.divide(weights, np.expand_dims(np.sum(weights, axis=1), axis=1))
    aux = np.concatenate((np.arange(in_length), np.arange(in_length - 1, -1, step=-1))).astype(np.int32)
    indices = aux[np.mod(indices, aux.size)]
    keep = np.nonzero(np.any(weights, axis=0))[0]
    return np.ascontiguousarray(weights[:, keep]), np.ascontiguousarray(indices[:, keep])


def _resize_along(img, weights, indices, dim):
    """imresize.py:50-74 for float inputs: out = sum_t w[:, t] * img[ind[:, t]] (sequential adds)."""
    img = np.moveaxis(img, dim, 0).astype(np.float64)
    acc = None
    for t in range(weights.shape[1]):
        w = weights[:, t].reshape((-1,) + (1,) * (img.ndim - 1))
        term = np.multiply(img[indices[:, t]], w)
        acc = term if acc is None else acc + term
    return np.moveaxis(acc, 0, dim)


def imresize(I, scalar_scale=None, output_shape=None):
    """Bicubic resize of (h, w[, C]) -> float64 (imresize.py:80-112); float inputs only."""
    I = np.asarray(I)
    if scalar_scale is not None:
        scale = [float(scalar_scale)] * 2
        output_size = [int(ceil(scale[k] * I.shape[k])) for k in range(2)]
    elif output_shape is not None:
        scale = [1.0 * output_shape[k] / I.shape[k] for k in range(2)]
        output_size = list(output_shape)
    else:
        print('Error: scalar_scale OR output_shape should be defined!')
        return None
    order = np.argsort(np.array(scale))
    B = I
    for k in order:
        w, ind = contributions(I.shape[k], output_size[k], scale[k])
        B = _resize_along(B, w, ind, int(k))
    return B
