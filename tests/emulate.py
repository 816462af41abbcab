"""CPU emulation of the roundings of the device path (tests only): which operand is fp16, where the trunk is re-coded.

Convolutions run in float64 on operands rounded exactly as the kernels round them, so what is measured is the error
budget of the precision DESIGN (fp16 operands, fp32-class accumulation, split first / last layer, trunk format), not an
implementation.  ``trunk``: 'q8' (fp16 + 8 bits, the inference path), 'fp32', or 'fp16' (no extra bits)."""
import numpy as np
import torch
import torch.nn.functional as F

from qtrunk import q_decode, q_encode


def _conv64(x, w_hwio, b):
    w = torch.from_numpy(np.ascontiguousarray(w_hwio.transpose(3, 2, 0, 1)).astype(np.float64))
    return F.conv2d(torch.from_numpy(x.astype(np.float64)), w, torch.from_numpy(b.astype(np.float64)), padding=1).numpy()


def _f16(a):
    return a.astype(np.float16).astype(np.float32)


def _split(a):
    hi = _f16(a)
    return hi, _f16(a - hi)


def forward_emulated(inputs, weights, trunk='q8', split_ends=True, scale=0.1):
    x_in = np.concatenate([np.asarray(a, np.float32) for a in inputs], axis=1)
    skip = np.asarray(inputs[-1], np.float32)
    w0, b0 = weights[0]
    if split_ends:                                            # hi + lo operands: fp32-equivalent first layer
        xh, xl = _split(x_in)
        wh, wl = _split(w0)
        acc = _conv64(xh, wh, b0) + _conv64(xl, wh, 0 * b0) + _conv64(xh, wl, 0 * b0) + _conv64(xl, wl, 0 * b0)
    else:
        acc = _conv64(_f16(x_in), _f16(w0), b0)
    x = np.maximum(acc, 0).astype(np.float32)

    def recode(v):                                            # -> (trunk value kept, fp16 operand of the next conv)
        if trunk == 'q8':
            h, lo = q_encode(v)
            return q_decode(h, lo), h.astype(np.float32)
        if trunk == 'fp16':
            return _f16(v), _f16(v)
        return v, _f16(v)

    x, x_hi = recode(x)
    n_res = (len(weights) - 2) // 2
    for l in range(n_res):
        w1, b1 = weights[1 + 2 * l]
        w2, b2 = weights[2 + 2 * l]
        t = _f16(np.maximum(_conv64(x_hi, _f16(w1), b1), 0).astype(np.float32))
        u = _conv64(t, _f16(w2), b2).astype(np.float32)
        x, x_hi = recode((x + np.float32(scale) * u).astype(np.float32))
    wt, bt = weights[-1]
    if split_ends:
        xh, xl = _split(x)
        wh, wl = _split(wt)
        out = _conv64(xh, wh, bt) + _conv64(xl, wh, 0 * bt) + _conv64(xh, wl, 0 * bt) + _conv64(xl, wl, 0 * bt)
    else:
        out = _conv64(_f16(x), _f16(wt), bt)
    return (out + skip).astype(np.float32)


def forward_exact(inputs, weights, scale=0.1):
    """float64 evaluation of the graph (DSen2Net.py:18-43) on the float32 weights."""
    x_in = np.concatenate([np.asarray(a, np.float64) for a in inputs], axis=1)
    x = np.maximum(_conv64(x_in, *weights[0]), 0)
    for l in range((len(weights) - 2) // 2):
        t = np.maximum(_conv64(x, *weights[1 + 2 * l]), 0)
        x = x + scale * _conv64(t, *weights[2 + 2 * l])
    return _conv64(x, *weights[-1]) + np.asarray(inputs[-1], np.float64)


# ---- feasibility study for the next round: Winograd F(2x2, 3x3) with fp16 transformed operands -----------------------
_BT = np.array([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], np.float64)
_G = np.array([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], np.float64)
_AT = np.array([[1, 1, 1, 0], [0, 1, -1, -1]], np.float64)


def conv_winograd_f16(x, w_hwio, b):
    """3x3 'same' convolution as Winograd F(2x2,3x3): 16 channel contractions (2.25x fewer multiplications) on the
    TRANSFORMED input tiles and kernels, both rounded to fp16 (what tensor-core operands would be); float64 sums."""
    n, C, H, W = x.shape
    assert H % 2 == 0 and W % 2 == 0
    xp = torch.from_numpy(np.pad(x.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1))))
    d = xp.unfold(2, 4, 2).unfold(3, 4, 2)                       # (n, C, H/2, W/2, 4, 4)
    bt = torch.from_numpy(_BT)
    v = torch.einsum('ij,nchwjk,lk->nchwil', bt, d, bt)          # B^T d B
    v = v.to(torch.float16).to(torch.float64)
    g = torch.from_numpy(w_hwio.astype(np.float64)).permute(3, 2, 0, 1)      # (Cout, Cin, 3, 3)
    gm = torch.from_numpy(_G)
    u = torch.einsum('ij,ocjk,lk->ocil', gm, g, gm)              # G g G^T
    u = u.to(torch.float16).to(torch.float64)
    m = torch.einsum('ocil,nchwil->nohwil', u, v)
    at = torch.from_numpy(_AT)
    y = torch.einsum('ij,nohwjk,lk->nohwil', at, m, at)          # (n, Cout, H/2, W/2, 2, 2)
    y = y.permute(0, 1, 2, 4, 3, 5).reshape(n, -1, H, W)
    return (y + torch.from_numpy(b.astype(np.float64)).view(1, -1, 1, 1)).numpy()


def forward_emulated_winograd(inputs, weights, scale=0.1):
    """forward_emulated(trunk='q8') with the 12 trunk convolutions in Winograd form."""
    x_in = np.concatenate([np.asarray(a, np.float32) for a in inputs], axis=1)
    skip = np.asarray(inputs[-1], np.float32)
    w0, b0 = weights[0]
    xh, xl = _split(x_in)
    wh, wl = _split(w0)
    acc = _conv64(xh, wh, b0) + _conv64(xl, wh, 0 * b0) + _conv64(xh, wl, 0 * b0) + _conv64(xl, wl, 0 * b0)
    h, lo = q_encode(np.maximum(acc, 0).astype(np.float32))
    x, x_hi = q_decode(h, lo), h.astype(np.float32)
    for l in range((len(weights) - 2) // 2):
        w1, b1 = weights[1 + 2 * l]
        w2, b2 = weights[2 + 2 * l]
        t = _f16(np.maximum(conv_winograd_f16(x_hi, w1, b1), 0).astype(np.float32))
        u = conv_winograd_f16(t, w2, b2).astype(np.float32)
        h, lo = q_encode((x + np.float32(scale) * u).astype(np.float32))
        x, x_hi = q_decode(h, lo), h.astype(np.float32)
    wt, bt = weights[-1]
    xh, xl = _split(x)
    wh, wl = _split(wt)
    out = _conv64(xh, wh, bt) + _conv64(xl, wh, 0 * bt) + _conv64(xh, wl, 0 * bt) + _conv64(xl, wl, 0 * bt)
    return (out + skip).astype(np.float32)
