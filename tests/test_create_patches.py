"""Host logic of the training-data generation front end (dsen2_b200/create_patches.py, create_random.py) against values worked
out by hand from training/create_patches.py:27-30,60-71,207-316 and create_random.py:11-19 on a synthetic .npz product.  The
GPU pieces it calls (downPixelAggr, the patch savers) have their own parity tests (tests/test_gpu_patches.py) and are stubbed
here."""
import json
import random
import struct
import zlib

import numpy as np
import pytest

from dsen2_b200 import create_patches as cp
from dsen2_b200 import create_random, patches
from test_s2_tiles import D10, D20, D60, _product


def test_roi_is_shrunk_to_36_pixel_boundaries():
    assert cp.clamp_roi_36(40, 50, 400, 500, 10980, 10980) == (36, 36, 395, 467)         # int(40/36)*36, int(401/36)*36 - 1
    assert cp.clamp_roi_36(400, 500, 40, 50, 10980, 10980) == (36, 36, 395, 467)         # corners in any order
    assert cp.clamp_roi_36(-5, 0, 20000, 71, 10980, 10980) == (0, 0, 10979, 71)          # clamped: 10980 = 305 * 36
    assert cp.clamp_roi_36(10, 10, 30, 30, 10980, 10980) == (0, 0, -1, -1)               # less than one block: empty


def test_band_selection_takes_the_fixed_lists():
    (n10, i10), (n20, i20), (n60, i60) = cp.select_bands(D10, D20, D60, run_60=False)
    assert n10 == ['B4', 'B3', 'B2', 'B8'] and i10 == [0, 1, 2, 3]
    assert n20 == ['B5', 'B6', 'B7', 'B8A', 'B11', 'B12'] and (n60, i60) == ([], [])
    assert cp.select_bands(D10, D20, D60, run_60=True)[2] == (['B1', 'B9'], [0, 1])      # B10 is never selected


def _png_pixels(path):
    raw = open(path, 'rb').read()
    assert raw[:8] == b'\x89PNG\r\n\x1a\n'
    pos, chunks = 8, {}
    while pos < len(raw):
        n, tag = struct.unpack('>I4s', raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        assert struct.unpack('>I', raw[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xffffffff
        chunks.setdefault(tag, b'')
        chunks[tag] += body
        pos += 12 + n
    w, h, depth, ctype = struct.unpack('>IIBB', chunks[b'IHDR'][:10])
    c = {0: 1, 2: 3}[ctype]
    data = np.frombuffer(zlib.decompress(chunks[b'IDAT']), np.uint8).reshape(h, 1 + w * c)
    assert depth == 8 and not data[:, 0].any()                                           # filter type 0 on every row
    return data[:, 1:].reshape(h, w, c)


def test_quick_look_png(tmp_path):
    rng = np.random.RandomState(1)
    img = rng.rand(5, 7, 3) * 4000
    cp.save_band(str(tmp_path) + '/', img, 'raw/rgbs/xRGB')
    mi, ma = np.percentile(img, (1, 99))
    want = ((np.clip(img, mi, ma) - mi) / (ma - mi) * 255 + 0.5).astype(np.uint8)
    assert np.array_equal(_png_pixels(str(tmp_path / 'raw' / 'rgbs' / 'xRGB.png')), want)


@pytest.fixture
def stubs(monkeypatch):
    calls = []

    def down(img, SCALE=2):
        calls.append(('down', img.shape, SCALE))
        return img[::SCALE, ::SCALE].astype(np.float64)

    def saver(name):
        def f(*a, **kw):
            calls.append((name, [x.shape for x in a[:-1]], a[-1], kw))
        return f
    monkeypatch.setattr(patches, 'downPixelAggr', down)
    for n in ('save_test_patches', 'save_test_patches60', 'save_random_patches', 'save_random_patches60'):
        monkeypatch.setattr(patches, n, saver(n))
    return calls


def test_training_patches_20m(tmp_path, stubs, capsys):
    """default mode (:303-316): both inputs degraded by 2, the 20 m bands at full resolution are the labels."""
    path = _product(tmp_path, H=144, W=216)
    prefix = str(tmp_path) + '/data/'
    (tmp_path / 'data').mkdir()
    assert cp.main([path, '--save_prefix', prefix, '--roi_x_y', '0,0,215,100']) == 0
    assert 'Selected pixel region: xmin=0, ymin=0, xmax=215, ymax=71' in capsys.readouterr().out     # int(101/36)*36 - 1
    assert stubs[0] == ('down', (72, 216, 4), 2) and stubs[1] == ('down', (36, 108, 6), 2)
    name, shapes, out, _ = stubs[2]
    assert name == 'save_random_patches' and shapes == [(36, 108, 6), (36, 108, 4), (18, 54, 6)]      # gt, 10 m lr, 20 m lr
    assert out == prefix + 'train/product.npz/'
    assert (tmp_path / 'data' / 'train' / 'product.npz').is_dir()


def test_test_patches_60m_and_true_scale(tmp_path, stubs):
    """--test_data --run_60 (:241-272): everything degraded by 6, roi.json in 60 m-degraded pixels, untiled copies;
    --true_data (:283-301): no degradation, 384 / 12 patches, roi.json in 10 m pixels."""
    path = _product(tmp_path, H=144, W=216)
    prefix = str(tmp_path) + '/data/'
    (tmp_path / 'data').mkdir()
    assert cp.main([path, '--save_prefix', prefix, '--test_data', '--run_60']) == 0
    assert [c[2] for c in stubs[:3]] == [6, 6, 6] and stubs[2][1] == (24, 36, 2)
    name, shapes, out, _ = stubs[3]
    assert name == 'save_test_patches60' and shapes == [(24, 36, 4), (12, 18, 6), (4, 6, 2)]
    d = tmp_path / 'data' / 'test60' / 'product.npz'
    assert json.load(open(d / 'roi.json')) == [0, 0, 36, 24]
    assert np.load(d / 'no_tiling' / 'data60_gt.npy').shape == (24, 36, 2)
    assert np.load(d / 'no_tiling' / 'data10.npy').dtype == np.float32
    stubs.clear()
    assert cp.main([path, '--save_prefix', prefix, '--true_data', '--run_60']) == 0
    name, shapes, out, kw = stubs[0]                                                     # no degradation at all
    assert name == 'save_test_patches60' and shapes == [(144, 216, 4), (72, 108, 6), (24, 36, 2)]
    assert kw == dict(patchSize=384, border=12)
    assert json.load(open(tmp_path / 'data' / 'true' / 'product.npz' / 'roi.json')) == [0, 0, 216, 144]
    with pytest.raises(SystemExit):
        cp.main([path, '--save_prefix', prefix, '--true_data'])                         # the 60 m bands are not selected


def test_test_patches_20m_write_the_quick_look(tmp_path, stubs):
    path = _product(tmp_path, H=144, W=216)
    prefix = str(tmp_path) + '/data/'
    (tmp_path / 'data').mkdir()
    assert cp.main([path, '--save_prefix', prefix, '--test_data']) == 0
    d = tmp_path / 'data' / 'test' / 'product.npz'
    assert json.load(open(d / 'roi.json')) == [0, 0, 108, 72]
    assert np.load(d / 'no_tiling' / 'data20_gt.npy').shape == (72, 108, 6)
    assert _png_pixels(str(d / 'RGB.png')).shape == (72, 108, 3)


def test_validation_index(tmp_path):
    random.seed(5)
    index, draws = create_random.make_val_index(1000, 0.1)
    assert index.dtype == bool and index.sum() == 100 and draws >= 100
    random.seed(5)                                                                       # the reference's loop, re-stated
    ref = np.zeros(1000, bool)
    while ref.sum() < 100:
        ref[random.randrange(0, 1000)] = True
    assert np.array_equal(index, ref)
    assert create_random.main(['--tiles', '2', '--patches_per_tile', '50', '--path', str(tmp_path) + '/']) == 0
    assert np.load(tmp_path / 'val_index.npy').sum() == 10
