"""Error budget of the precision design, on the CPU (no GPU needed): the roundings of the device path applied to a
float64 evaluation of the DSen2 graph.  North-star gate: max |err| <= 5e-3 on the /2000-scaled output."""
import numpy as np

from emulate import forward_emulated, forward_exact
from oracle import dsen2net_oracle as no


def _case(seed=3, n=2, P=40):
    rng = np.random.RandomState(seed)
    xs = [(0.8 + 0.45 * rng.randn(n, c, P, P)).clip(0, 6).astype(np.float32) for c in (4, 6)]
    w = no.he_uniform_weights(10, 6, 6, 128, seed=seed)
    w = [(k, (0.05 * rng.randn(*b.shape)).astype(np.float32)) for k, b in w]
    return xs, w


def test_trunk_formats_against_the_exact_graph():
    xs, w = _case()
    ref = forward_exact(xs, w)
    err = {t: np.abs(forward_emulated(xs, w, trunk=t) - ref).max() for t in ('fp32', 'q8', 'fp16')}
    print('max abs error vs float64:', {k: float(v) for k, v in err.items()})
    assert err['fp32'] < 2.5e-3 and err['q8'] < 2.5e-3            # gate 5e-3
    assert abs(err['q8'] - err['fp32']) < 1e-4                     # 19 significant bits: indistinguishable from fp32
    assert err['fp16'] > err['q8']                                 # an fp16 trunk is what the extra byte buys back


def test_split_first_and_last_layer_matter():
    xs, w = _case(seed=5)
    ref = forward_exact(xs, w)
    e_split = np.abs(forward_emulated(xs, w, trunk='q8', split_ends=True) - ref).max()
    e_plain = np.abs(forward_emulated(xs, w, trunk='q8', split_ends=False) - ref).max()
    print('split ends %.2e, fp16 ends %.2e' % (e_split, e_plain))
    assert e_split < e_plain


def test_winograd_trunk_feasibility():
    """Next-round study: Winograd F(2x2,3x3) trunk convolutions with fp16 transformed operands stay inside the gate
    (the transform itself is exact: checked against the direct convolution in float64 first)."""
    from emulate import conv_winograd_f16, forward_emulated_winograd, _conv64
    rng = np.random.RandomState(1)
    x = rng.randn(1, 8, 12, 12).astype(np.float16).astype(np.float32)
    w = (rng.randn(3, 3, 8, 5) * 0.1).astype(np.float32)
    b = rng.randn(5).astype(np.float32)
    np.testing.assert_allclose(conv_winograd_f16(x, w, b), _conv64(x, w, b), rtol=0, atol=5e-3)   # fp16 operand rounding only
    xs, wts = _case()
    ref = forward_exact(xs, wts)
    e_direct = np.abs(forward_emulated(xs, wts, trunk='q8') - ref).max()
    e_wino = np.abs(forward_emulated_winograd(xs, wts) - ref).max()
    print('direct %.2e, winograd %.2e' % (e_direct, e_wino))
    assert e_wino < 5e-3


def test_vdsen2_depth_32_budget():
    """The same budget at VDSen2's real depth (32 resBlocks x 256 features, he_uniform weights as in the GPU test and the
    bench): the error of the built design is the fp16 rounding of the 64 trunk convolutions' operands -- the fp16 + 8 bit
    trunk is indistinguishable from an fp32 trunk -- and an fp16-only trunk would break the 5e-3 gate."""
    rng = np.random.RandomState(3)
    xs = [(0.8 + 0.45 * rng.randn(1, c, 48, 48)).clip(0, 6).astype(np.float32) for c in (4, 6)]
    w = no.he_uniform_weights(10, 6, 32, 256, seed=0)
    ref = forward_exact(xs, w)
    err = {t: float(np.abs(forward_emulated(xs, w, trunk=t) - ref).max()) for t in ('fp32', 'q8', 'fp16')}
    print('VDSen2 32 x 256, outputs up to %.1f: max abs error vs float64' % np.abs(ref).max(), err)
    assert err['q8'] < 5e-3 and err['fp32'] < 5e-3
    assert abs(err['q8'] - err['fp32']) < 5e-4
    assert err['fp16'] > 5e-3
