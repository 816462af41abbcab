"""CUDA patch / resize / stitch kernels against the CPU oracle and the reference fingerprints (bit-exact)."""
import numpy as np
import pytest
from conftest import sha16

from cases import CASES20, CASES60, synth, synth_pred

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mods():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from dsen2_b200 import imresize, patches
    from oracle import imresize_oracle, patches_oracle
    patches.VERBOSE = False
    return patches, imresize, patches_oracle, imresize_oracle


def test_extract_scene_bit_exact(mods, malmo, fingerprints):
    patches, _, po, _ = mods
    d10, d20, d60 = malmo
    p10, p20 = patches.get_test_patches(d10, d20, 128, 8, interp=False)
    assert sha16(p10) == fingerprints['malmo.p10'] and sha16(p20) == fingerprints['malmo.p20']
    q10, q20, q60 = patches.get_test_patches60(d10, d20, d60, 192, 12, interp=False)
    assert (sha16(q10), sha16(q20), sha16(q60)) == tuple(fingerprints['malmo.' + k] for k in ('q10', 'q20', 'q60'))
    # uint16 input (what GDAL hands s2_tiles_supres.py) gives the same stacks
    r10, r20 = patches.get_test_patches(d10.astype(np.uint16), d20.astype(np.uint16), 128, 8, interp=False)
    assert np.array_equal(r10, p10) and np.array_equal(r20, p20)


@pytest.mark.parametrize('tag', sorted(CASES20))
def test_extract_stitch_20_synthetic(mods, tag, fingerprints):
    patches, _, po, _ = mods
    d10, d20, _ = synth(tag)
    p10, p20 = patches.get_test_patches(d10, d20, 128, 8, interp=False)
    assert p10.shape[0] == fingerprints['syn_%s.n' % tag]
    assert sha16(p10) == fingerprints['syn_%s.p10' % tag] and sha16(p20) == fingerprints['syn_%s.p20' % tag]
    pred = synth_pred(tag, p10.shape[0], 3, 128)
    rec = patches.recompose_images(pred, 8, d10.shape)
    assert rec.shape == d10.shape[:2] + (3,) and rec.dtype == np.float32
    assert sha16(rec) == fingerprints['syn_%s.rec' % tag]
    assert np.array_equal(patches.recompose_images(p10, 8, d10.shape), d10)


@pytest.mark.parametrize('tag', sorted(CASES60))
def test_extract_stitch_60_synthetic(mods, tag, fingerprints):
    patches, _, po, _ = mods
    d10, d20, d60 = synth(tag)
    q10, q20, q60 = patches.get_test_patches60(d10, d20, d60, 192, 12, interp=False)
    assert (sha16(q10), sha16(q20), sha16(q60)) == tuple(fingerprints['syn_%s.%s' % (tag, k)] for k in ('q10', 'q20', 'q60'))
    rec = patches.recompose_images(synth_pred(tag, q10.shape[0], 2, 192), 12, d10.shape)
    assert sha16(rec) == fingerprints['syn_%s.rec' % tag]
    assert np.array_equal(patches.recompose_images(q10, 12, d10.shape), d10)


def test_extract_shard_ranges_and_divisor(mods):
    import torch
    patches, _, po, _ = mods
    d10, d20, _ = synth('a')
    ref10, ref20 = po.get_test_patches(d10, d20, 128, 8, interp=False)
    t10 = torch.from_numpy(d10).cuda()
    n = ref10.shape[0]
    parts = [patches.extract_patches_device(t10, 2, 64, 4, a, b - a).cpu().numpy()
             for a, b in ((0, 5), (5, 6), (6, n))]
    assert np.array_equal(np.concatenate(parts), ref10)
    scaled = patches.extract_patches_device(t10, 2, 64, 4, divisor=2000.0).cpu().numpy()
    assert np.array_equal(scaled, ref10 / np.float32(2000))          # IEEE division, bit-exact (supres.py:23)
    with pytest.raises(Exception):
        patches.extract_patches_device(t10, 2, 64, 4, n - 1, 5)      # beyond the allocated stack


def test_recompose_sharded_matches_sequential_overwrite(mods):
    import torch
    patches, _, po, _ = mods
    d10, _, _ = synth('a')                                          # 300x412: clamped last row and column
    H, W = d10.shape[:2]
    pred = synth_pred('a', 12, 6, 128)
    ref = po.recompose_images(pred, 8, (H, W))
    out = torch.zeros((H, W, 6), device='cuda')
    tp = torch.from_numpy(pred).cuda()
    for a, b in ((7, 12), (0, 3), (3, 7)):                          # any order: ownership, not overwrite order
        patches.recompose_device(tp[a:b].contiguous(), 8, H, W, first_patch=a, mul=1.0, out=out)
    assert np.array_equal(out.cpu().numpy(), ref)
    out2 = patches.recompose_device(tp, 8, H, W, mul=2000.0).cpu().numpy()
    assert np.array_equal(out2, ref * np.float32(2000))


def test_recompose_single_patch_and_errors(mods):
    patches, _, po, _ = mods
    a = np.random.RandomState(0).rand(1, 3, 32, 32).astype(np.float32)
    assert np.array_equal(patches.recompose_images(a, 4, (24, 24)), po.recompose_images(a, 4, (24, 24)))
    with pytest.raises(ValueError):
        patches.recompose_images(np.zeros((4, 3, 128, 128), np.float32), 8, (100, 300))   # H < interior


def test_bilinear_matches_oracle(mods, malmo):
    patches, _, po, _ = mods
    d10, d20, d60 = malmo
    p10, p20 = po.get_test_patches(d10, d20, 128, 8, interp=False)
    got = patches.interp_patches(p20, p10.shape)
    ref = po.interp_patches(p20, p10.shape)
    assert got.shape == ref.shape == (36, 6, 128, 128) and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-3)          # DN units; fp32 vs the oracle's f64 interpolation
    q10, q20, q60 = po.get_test_patches60(d10, d20, d60, 192, 12, interp=False)
    np.testing.assert_allclose(patches.interp_patches(q60, q10.shape), po.interp_patches(q60, q10.shape), rtol=0, atol=2e-3)
    np.testing.assert_allclose(patches.interp_patches(q20, q10.shape), po.interp_patches(q20, q10.shape), rtol=0, atol=2e-3)
    full10, full20 = patches.get_test_patches(d10, d20, 128, 8)       # interp=True default path
    np.testing.assert_allclose(full20, ref, rtol=0, atol=2e-3)
    assert np.array_equal(full10, p10)


def test_bicubic_bit_identical(mods, malmo, golden, fingerprints):
    _, imresize, _, io_ = mods
    _, d20, d60 = malmo
    b2 = imresize.imresize(d20, 2)
    assert b2.dtype == np.float64 and b2.shape == (600, 600, 6)
    assert sha16(b2) == fingerprints['malmo.bic2']
    b6 = imresize.imresize(d60, 6)
    assert sha16(b6) == fingerprints['malmo.bic6']
    img = golden['bic_in']
    assert np.array_equal(imresize.imresize(img, 2), golden['bic_out2'])
    assert np.array_equal(imresize.imresize(img, 6), golden['bic_out6'])
    assert np.array_equal(imresize.imresize(img, output_shape=(20, 11)), golden['bic_out_shape'])
    assert np.array_equal(imresize.imresize(img[:, :, 0], 2), golden['bic_out2'][:, :, 0])     # 2-D input
    x64 = np.random.RandomState(1).rand(17, 23, 3) * 1e4
    assert np.array_equal(imresize.imresize(x64, 3), io_.imresize(x64, 3))                      # float64 input
    assert imresize.imresize(img) is None                                                       # imresize.py:91-93


def test_full_tile_roundtrip_property(mods):
    """BASELINE full-size geometry (10980^2) through a size-independent property: stitch(extract(x)) == x."""
    import torch
    patches, _, _, _ = mods
    H = W = 10980
    g = torch.Generator(device='cuda').manual_seed(20170928)
    d10 = torch.randint(0, 12000, (H, W, 4), generator=g, device='cuda').float()
    alloc, filled = patches.patch_counts(H // 2, W // 2, 64, 4)
    assert (alloc, filled) == (9801, 9801)
    out = torch.zeros((H, W, 4), device='cuda')
    for p0 in range(0, filled, 1024):
        nb = min(1024, filled - p0)
        p10 = patches.extract_patches_device(d10, 2, 64, 4, p0, nb)
        patches.recompose_device(p10, 8, H, W, first_patch=p0, out=out)
    assert torch.equal(out, d10)
    # 60 m geometry: 66x66 patches, last two starts 1792 / 1802 on the 60 m grid
    alloc60, filled60 = patches.patch_counts(1830, 1830, 32, 2)
    assert (alloc60, filled60) == (4356, 4356)
    out.zero_()
    for p0 in range(0, filled60, 512):
        nb = min(512, filled60 - p0)
        q10 = patches.extract_patches_device(d10, 6, 32, 2, p0, nb)
        patches.recompose_device(q10, 12, H, W, first_patch=p0, out=out)
    assert torch.equal(out, d10)


@pytest.mark.parametrize('shape,scale', [((48, 60, 4), 2), ((48, 60, 6), 6), ((36, 36, 2), 6), ((20, 30, 1), 2)])
def test_down_pixel_aggr_matches_scipy_oracle(mods, shape, scale):
    """patches.py:353-371: gaussian_filter(sigma = 1/scale) + block mean; oracle = scipy's gaussian_filter itself."""
    patches, _, po, _ = mods
    rng = np.random.RandomState(shape[0] + scale)
    img = (rng.rand(*shape) * 4000).astype(np.float32)
    got = patches.downPixelAggr(img, SCALE=scale)
    ref = po.downPixelAggr(img, SCALE=scale)
    assert got.shape == ref.shape and got.dtype == np.float64
    print('downPixelAggr max |diff| =', np.abs(got - ref).max())
    assert np.array_equal(got, ref)       # same double arithmetic, float32 storage between the passes, same summation order
    got2 = patches.downPixelAggr(img[:, :, 0], SCALE=scale)         # 2-D input is expanded and squeezed again
    assert np.array_equal(got2, ref[..., 0] if ref.ndim == 3 else ref)
    # uint16 digital numbers (what create_patches.py reads through GDAL): scipy stores each pass as uint16 = truncation
    dn = (rng.rand(*shape) * 9000).astype(np.uint16)
    dn[:4, :4] = 1234                                               # a flat patch: sums that land on an integer boundary
    ref16 = po.downPixelAggr(dn, SCALE=scale)
    assert np.array_equal(patches.downPixelAggr(dn, SCALE=scale), ref16)
    assert not np.array_equal(ref16, po.downPixelAggr(dn.astype(np.float32), SCALE=scale))


def test_training_patch_writers(mods, tmp_path):
    """save_test_patches / save_random_patches (patches.py:159-224) write the stacks supres_train.py loads."""
    import random
    patches, _, po, _ = mods
    rng = np.random.RandomState(4)
    gt20 = (rng.rand(96, 120, 6) * 3000).astype(np.float32)
    d10 = (rng.rand(96, 120, 4) * 3000).astype(np.float32)
    d20 = (rng.rand(48, 60, 6) * 3000).astype(np.float32)
    pre = str(tmp_path) + '/'
    random.seed(11)
    patches.save_random_patches(gt20, d10, d20, pre, NR_CROP=7)
    a10, a20, agt = (np.load(pre + k + '.npy') for k in ('data10', 'data20', 'data20_gt'))
    assert a10.shape == (7, 4, 32, 32) and a20.shape == (7, 6, 32, 32) and agt.shape == (7, 6, 32, 32)
    random.seed(11)                                                     # replay the crops with the oracle's upsampling
    for i in range(7):
        x0, y0 = random.randrange(0, 48 - 16), random.randrange(0, 60 - 16)
        assert np.array_equal(a10[i], np.rollaxis(d10[2 * x0:2 * x0 + 32, 2 * y0:2 * y0 + 32], 2))
        assert np.array_equal(agt[i], np.rollaxis(gt20[2 * x0:2 * x0 + 32, 2 * y0:2 * y0 + 32], 2))
        lr = np.rollaxis(d20[x0:x0 + 16, y0:y0 + 16], 2)[None]
        np.testing.assert_allclose(a20[i], po.interp_patches(lr, (1, 4, 32, 32))[0], rtol=0, atol=2e-3)
    patches.save_test_patches(d10, d20, pre + 't_', patchSize=32, border=4)
    t10, t20 = np.load(pre + 't_data10.npy'), np.load(pre + 't_data20.npy')
    r10, r20 = po.get_test_patches(d10, d20, 32, 4)
    assert np.array_equal(t10, r10)
    np.testing.assert_allclose(t20, r20, rtol=0, atol=2e-3)


def test_create_patches_cli_on_an_npz_product(mods, tmp_path):
    """training/create_patches.py end to end on the GPU kernels (degradation, tiling, bilinear upsampling), .npz product:
    test patches == the oracle's get_test_patches of the oracle-degraded images; random training crops have the documented
    shapes and the label is the full-resolution 20 m crop."""
    _, _, po, _ = mods
    from dsen2_b200 import create_patches as cp
    rng = np.random.RandomState(4)
    H, W = 648, 720                  # 60 m after the 6x degradation: 18 x 20 pixels, enough for the 16 x 16 random crops
    d10 = rng.randint(1, 9000, (H, W, 4)).astype(np.uint16)
    d20 = rng.randint(1, 9000, (H // 2, W // 2, 6)).astype(np.uint16)
    d60 = rng.randint(1, 9000, (H // 6, W // 6, 2)).astype(np.uint16)
    desc = lambda names: np.array(["%s, central wavelength 500 nm" % n for n in names])
    path = str(tmp_path / 'S2A_TEST.npz')
    np.savez(path, data10=d10, data20=d20, data60=d60, desc10=desc(['B4', 'B3', 'B2', 'B8']),
             desc20=desc(['B5', 'B6', 'B7', 'B8A', 'B11', 'B12']), desc60=desc(['B1', 'B9']),
             geotransform=np.array([3e5, 10., 0., 5e6, 0., -10.]), utm=np.array('UTM 32N'))
    prefix = str(tmp_path) + '/data/'
    (tmp_path / 'data').mkdir()
    assert cp.main([path, '--save_prefix', prefix, '--test_data']) == 0
    d = tmp_path / 'data' / 'test' / 'S2A_TEST.npz'
    lr10, lr20 = po.downPixelAggr(d10, 2), po.downPixelAggr(d20, 2)
    assert np.array_equal(np.load(d / 'no_tiling' / 'data10.npy'), lr10.astype(np.float32))          # bit-identical to scipy
    p10, p20 = po.get_test_patches(lr10, lr20, 128, 4)
    assert np.array_equal(np.load(d / 'data10.npy'), p10)
    np.testing.assert_allclose(np.load(d / 'data20.npy'), p20, rtol=0, atol=2e-3)
    assert cp.main([path, '--save_prefix', prefix, '--run_60']) == 0                                  # 500 random crops
    t = tmp_path / 'data' / 'train60' / 'S2A_TEST.npz'
    assert np.load(t / 'data10.npy').shape == (500, 4, 96, 96) and np.load(t / 'data60.npy').shape == (500, 2, 96, 96)
    assert np.load(t / 'data60_gt.npy').shape == (500, 2, 96, 96)
