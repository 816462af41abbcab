"""The CPU oracle against the reference's own outputs (golden fixtures made by tests/golden/make_golden.py)."""
import numpy as np
import pytest
from conftest import sha16

from cases import CASES20, CASES60, synth, synth_pred
from oracle import imresize_oracle as io_
from oracle import patches_oracle as po


def test_scene_fixture_matches_reference_fingerprint(scene, fingerprints):
    import hashlib
    name, d10, d20, d60 = scene
    for a, k in ((d10, 'im10'), (d20, 'im20'), (d60, 'im60')):
        h5order = np.ascontiguousarray(a.transpose())  # readh5 applies .transpose() (demoDSen2.py:16-23)
        assert hashlib.sha1(h5order.tobytes()).hexdigest()[:12] == fingerprints['%s.%s' % (name, k)]


def test_extract_scene_matches_reference(scene, fingerprints):
    name, d10, d20, d60 = scene
    p10, p20 = po.get_test_patches(d10, d20, 128, 8, interp=False)
    assert p10.shape == (36, 4, 128, 128) and p20.shape == (36, 6, 64, 64)
    assert sha16(p10) == fingerprints[name + '.p10'] and sha16(p20) == fingerprints[name + '.p20']
    q10, q20, q60 = po.get_test_patches60(d10, d20, d60, 192, 12, interp=False)
    assert q10.shape == (16, 4, 192, 192) and q60.shape == (16, 2, 32, 32)
    assert (sha16(q10), sha16(q20), sha16(q60)) == tuple(fingerprints['%s.%s' % (name, k)] for k in ('q10', 'q20', 'q60'))


def test_bicubic_scene_matches_reference(shark, fingerprints):
    """imresize of the second scene (the Malmo one is covered below): sha1 of the float64 output of the reference's own run."""
    _, d20, d60 = shark
    assert sha16(io_.imresize(d20, 2)) == fingerprints['shark.bic2']
    assert sha16(io_.imresize(d60, 6)) == fingerprints['shark.bic6']


@pytest.mark.parametrize('tag', sorted(CASES20))
def test_extract_and_stitch_20_synthetic(tag, fingerprints):
    d10, d20, _ = synth(tag)
    p10, p20 = po.get_test_patches(d10, d20, 128, 8, interp=False)
    assert p10.shape[0] == fingerprints['syn_%s.n' % tag]
    assert sha16(p10) == fingerprints['syn_%s.p10' % tag]
    assert sha16(p20) == fingerprints['syn_%s.p20' % tag]
    rec = po.recompose_images(synth_pred(tag, p10.shape[0], 3, 128), 8, d10.shape)
    assert rec.shape == d10.shape[:2] + (3,) and rec.dtype == np.float32
    assert sha16(rec) == fingerprints['syn_%s.rec' % tag]
    # SURVEY section 4 invariant: stitching the un-interpolated 10 m patches returns the image
    assert np.array_equal(po.recompose_images(p10, 8, d10.shape), d10)


@pytest.mark.parametrize('tag', sorted(CASES60))
def test_extract_and_stitch_60_synthetic(tag, fingerprints):
    d10, d20, d60 = synth(tag)
    q10, q20, q60 = po.get_test_patches60(d10, d20, d60, 192, 12, interp=False)
    assert q10.shape[0] == fingerprints['syn_%s.n' % tag]
    assert (sha16(q10), sha16(q20), sha16(q60)) == tuple(fingerprints['syn_%s.%s' % (tag, k)] for k in ('q10', 'q20', 'q60'))
    rec = po.recompose_images(synth_pred(tag, q10.shape[0], 2, 192), 12, d10.shape)
    assert sha16(rec) == fingerprints['syn_%s.rec' % tag]
    assert np.array_equal(po.recompose_images(q10, 12, d10.shape), d10)


def test_surplus_zero_patches_when_divisible():
    # 560x560 allocates 36 patches, fills 25, the rest stay zero
    d10, d20, _ = synth('b')
    p10, p20 = po.get_test_patches(d10, d20, 128, 8, interp=False)
    assert p10.shape[0] == 36
    assert np.count_nonzero(p10[25:]) == 0 and np.count_nonzero(p20[25:]) == 0
    assert np.count_nonzero(p10[24]) > 0


def test_recompose_single_patch_returns_uncropped():
    a = np.random.RandomState(0).rand(1, 3, 32, 32).astype(np.float32)
    out = po.recompose_images(a, 4, (24, 24))
    assert out.shape == (32, 32, 3) and np.array_equal(out, a[0].transpose(1, 2, 0))


def test_bicubic_bit_identical_to_reference(golden, malmo, fingerprints):
    img = golden['bic_in']
    for k, kw in (('bic_out2', dict(scalar_scale=2)), ('bic_out6', dict(scalar_scale=6)),
                  ('bic_out_shape', dict(output_shape=(20, 11)))):
        out = io_.imresize(img, **kw)
        assert out.dtype == np.float64 and np.array_equal(out, golden[k]), k
    _, d20, d60 = malmo
    b2 = io_.imresize(d20, 2)
    assert sha16(b2) == fingerprints['malmo.bic2'] and abs(b2.mean() - fingerprints['malmo.bic2.mean']) < 1e-9
    assert np.array_equal(b2[:40, :40], golden['malmo_bic2_crop'])
    b6 = io_.imresize(d60, 6)
    assert sha16(b6) == fingerprints['malmo.bic6']
    assert np.array_equal(b6[-40:, -40:], golden['malmo_bic6_crop'])


def test_bicubic_known_phase_weights():
    # SURVEY 8(a) a9 known answers
    w, ind = io_.contributions(300, 600, 2.0)
    assert w.shape == (600, 4)
    np.testing.assert_allclose(w[0], [-0.0234375, 0.2265625, 0.8671875, -0.0703125], atol=1e-12)
    np.testing.assert_allclose(w[1], [-0.0703125, 0.8671875, 0.2265625, -0.0234375], atol=1e-12)
    assert list(ind[0]) == [1, 0, 0, 1]
    w6, _ = io_.contributions(100, 600, 6.0)
    np.testing.assert_allclose(w6[0], [-0.050637, 0.447049, 0.674479, -0.070891], atol=1e-6)


def test_bilinear_matches_scipy_zoom():
    from scipy import ndimage
    rng = np.random.RandomState(3)
    for p, s in ((64, 2), (32, 6), (96, 2)):
        x = rng.randint(0, 12000, size=(2, 3, p, p)).astype(np.float32)
        got = po.interp_patches(x, (2, 3, p * s, p * s))
        assert got.dtype == np.float32 and got.shape == (2, 3, p * s, p * s)
        for n in range(2):
            for c in range(3):
                ref = ndimage.zoom(x[n, c] / np.float32(30000), s, order=1, mode='mirror', grid_mode=True) * np.float32(30000)
                np.testing.assert_allclose(got[n, c], ref, rtol=0, atol=2e-3)
    # s=2 interior weights are {.25,.75}; mirror boundary: first output = in[0]*.75 + in[1]*.25
    x = np.zeros((1, 1, 4, 4), np.float32)
    x[0, 0, :, 0], x[0, 0, :, 1] = 3000, 6000
    np.testing.assert_allclose(po.interp_patches(x, (1, 1, 8, 8))[0, 0, 0, :2], [3750, 3750], atol=1e-3)


def test_dsen2net_oracle_shapes_and_params():
    from oracle import dsen2net_oracle as no
    w = no.he_uniform_weights(10, 6, 6, 128, seed=0)
    assert len(w) == 14 and sum(k.size + b.size for k, b in w) == 1789574
    w60 = no.he_uniform_weights(12, 2, 6, 128, seed=0)
    assert sum(k.size + b.size for k, b in w60) == 1787266
    small = no.he_uniform_weights(10, 6, 1, 16, seed=1)
    rng = np.random.RandomState(0)
    x10, x20 = rng.rand(2, 4, 16, 16).astype(np.float32), rng.rand(2, 6, 16, 16).astype(np.float32)
    y = no.forward([x10, x20], small)
    assert y.shape == (2, 6, 16, 16) and y.dtype == np.float32
    # zero tail kernel => output is exactly the global skip (DSen2Net.py:41)
    small[-1] = (np.zeros_like(small[-1][0]), small[-1][1])
    assert np.array_equal(no.forward([x10, x20], small), x20)


def test_dsen2net_oracle_against_an_independent_float64_direct_convolution():
    """The CNN oracle is unpinned by any reference artefact (Keras and the weight files are absent), so at least pin its
    restatement against a second, independent one: plain numpy float64, no torch -- zero 'same' padding, cross-correlation
    written as nine shifted einsum products over HWIO kernels (DSen2Net.py:9-43 with Keras' conventions).  A layout slip
    in the torch version (the HWIO -> OIHW permute, channel concatenation order, skip source) cannot hide behind this."""
    from oracle import dsen2net_oracle as no

    def conv(x, k, b):                                   # x (N,C,H,W), k (3,3,Cin,Cout) HWIO, b (Cout,)
        n, c, h, w = x.shape
        xp = np.zeros((n, c, h + 2, w + 2))
        xp[:, :, 1:-1, 1:-1] = x
        y = np.zeros((n, k.shape[3], h, w))
        for dy in range(3):
            for dx in range(3):
                y += np.einsum('nchw,co->nohw', xp[:, :, dy:dy + h, dx:dx + w], k[dy, dx].astype(np.float64))
        return y + b.astype(np.float64)[None, :, None, None]

    rng = np.random.RandomState(4)
    for chans, L, F in (((4, 6), 3, 16), ((4, 6, 2), 2, 24)):
        w = no.he_uniform_weights(sum(chans), chans[-1], L, F, seed=5)
        w = [(k, (rng.randn(*b.shape) * 0.1).astype(np.float32)) for k, b in w]       # non-zero biases
        xs = [rng.rand(2, c, 12, 10).astype(np.float32) * 3 for c in chans]
        x = np.concatenate([a.astype(np.float64) for a in xs], axis=1)                 # DSen2Net.py:24,26
        x = np.maximum(conv(x, *w[0]), 0)                                              # :29
        for l in range(L):                                                             # :31-32, resBlock :9-15
            t = np.maximum(conv(x, *w[1 + 2 * l]), 0)
            x = x + 0.1 * conv(t, *w[2 + 2 * l])
        ref = conv(x, *w[-1]) + xs[-1]                                                 # :35, :38 / :41
        got = no.forward(xs, w)
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-5)


def test_down_pixel_aggr_oracle_against_an_independent_formula():
    """oracle.downPixelAggr (scipy gaussian_filter + block mean, patches.py:353-371) vs a direct separable convolution
    with symmetric padding -- pins the restatement of block_reduce and the kernel radius / boundary convention."""
    from oracle import patches_oracle as po
    rng = np.random.RandomState(8)
    img = rng.rand(24, 30, 2) * 1000
    for s in (2, 6):
        sigma = 1.0 / s
        r = int(4.0 * sigma + 0.5)
        x = np.arange(-r, r + 1)
        w = np.exp(-0.5 * (x / sigma) ** 2)
        w /= w.sum()
        pad = np.pad(img, ((r, r), (r, r), (0, 0)), mode='symmetric')
        blur = sum(w[i] * pad[i:i + 24, :, :] for i in range(2 * r + 1))
        blur = sum(w[j] * blur[:, j:j + 30, :] for j in range(2 * r + 1))
        ref = blur.reshape(24 // s, s, 30 // s, s, 2).mean(axis=(1, 3))
        np.testing.assert_allclose(po.downPixelAggr(img, s), ref, rtol=1e-12, atol=1e-9)
