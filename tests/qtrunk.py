"""numpy statement of the fp16 + 8 bit trunk code of include/dsen2_b200.h (dsen2_conv_head16_q / dsen2_conv_resq)."""
import numpy as np

NORMAL = np.float32(2.0 ** -14)      # below this x_hi is an fp16 subnormal and the code keeps x to an absolute 2^-24 only


def _f16_toward_zero(v):
    """float32 -> float16, rounding toward zero (cvt.rz.f16.f32)."""
    h = v.astype(np.float16)
    over = np.abs(h.astype(np.float32)) > np.abs(v)
    bits = h.view(np.uint16).copy()
    bits[over] -= 1                    # one step toward zero in sign-magnitude
    return bits.view(np.float16)


def q_encode(x):
    """float32 (..., F) -> (x_hi float16, lo int8)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    s = x.view(np.uint32) + np.uint32(0x1010)
    with np.errstate(over='ignore'):
        h = _f16_toward_zero(s.view(np.float32))
    lo = (((s >> np.uint32(5)) & np.uint32(0xFF)).astype(np.int16) - 128).astype(np.int8)
    lo[h == 0] = 0
    return h, lo


def q_decode(h, lo):
    return (h.astype(np.float32).view(np.uint32) + (lo.astype(np.int32) << 5).astype(np.uint32)).view(np.float32)


def q_to_tiles(lo):        # NHWC int8 -> (n, H, W/8, F/16, 8, 16)
    n, H, W, F = lo.shape
    return np.ascontiguousarray(lo.reshape(n, H, W // 8, 8, F // 16, 16).transpose(0, 1, 2, 4, 3, 5))


def q_from_tiles(q):
    n, H, tx, c16, _, _ = q.shape
    return np.ascontiguousarray(q.transpose(0, 1, 2, 4, 3, 5)).reshape(n, H, tx * 8, c16 * 16)
